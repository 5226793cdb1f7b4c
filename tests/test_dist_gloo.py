"""The N > 1 host path (row-tile work stealing + gather) on CPU: world_size 2, gloo backend, with
the CPU oracle standing in for the per-rank renderer.  What is under test is raingun_b200/dist.py:
tile claiming, row lists, the P2P gather and the reassembled frame."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raingun_b200.dist import guided_chunks, hybrid_plan, n_tiles, render_frame_sharded, rows_of_tiles, static_chunk

W, H, TILE = 96, 54, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, schedule, out_dir, gather_mode="p2p", workers=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from raingun_b200.synth import make_scene

    data, _ = make_scene("C4", spheres=40, depth=4)
    O = oracle()
    calls = []

    def render_rowlist(rows, out):
        # renders each listed row with the oracle; rows are arbitrary, so go row by row
        buf = np.concatenate([O.render_rows(data, W, H, int(y), int(y) + 1, threads=1)[0] for y in rows], axis=0)
        out.copy_(torch.from_numpy(buf.reshape(-1)))
        calls.append(len(rows))
        return len(rows)

    host_frames = host_barrier = None
    if gather_mode == "host":   # one shared host frame every rank writes its rows into (pinning needs CUDA: off here)
        import ctypes

        from raingun_b200.dist import HostBarrier, SharedHostFrame
        host_frames = SharedHostFrame(W, H, rank, world, pin=False)
        host_barrier = HostBarrier(rank, world)   # the shared-memory end-of-frame barrier instead of a collective

        def render_rowlist(rows, frame_ptr):   # noqa: F811 - the "host" flavour: row y goes to frame_ptr + y * W * 4
            frame = np.ctypeslib.as_array(ctypes.cast(frame_ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(H, W, 4))
            for y in rows:
                frame[int(y)] = O.render_rows(data, W, H, int(y), int(y) + 1, threads=1)[0][0]
            calls.append(len(rows))
            return len(rows)

    tiles = []
    # frame ids repeat on purpose: a restarted render loop reuses them, and the work-stealing counter keys
    # must not (a reused key is already past its last chunk: the stealable tiles would silently stay unrendered)
    for it, frame_id in enumerate((1, 2, 1, 1)):
        res = render_frame_sharded(render_rowlist if workers == 1 else [render_rowlist] * workers, W, H, rank, world,
                                   frame_id, torch.device("cpu"),
                                   tile_rows=TILE, schedule=schedule, gather_mode=gather_mode, peer_frames=host_frames,
                                   host_barrier=host_barrier)
        tiles.append(sorted(res.my_tiles))
        if rank == 0:
            np.save(os.path.join(out_dir, f"frame{it}.npy"), np.array(res.frame.numpy()))
        if host_frames is not None:
            dist.barrier()           # rank 0 has copied the frame out before anyone overwrites it
            host_frames.array(frame_id)[:] = 0
            dist.barrier()
    if host_frames is not None:
        host_frames.close()
        host_barrier.close()
    np.save(os.path.join(out_dir, f"tiles{rank}.npy"), np.array(tiles[0], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("schedule,gather_mode,workers", [("steal", "p2p", 1), ("static", "p2p", 1), ("steal", "reduce", 1),
                                                          ("steal", "reduce", 2), ("static", "p2p", 3), ("steal", "host", 1),
                                                          ("static", "host", 2)])
def test_two_ranks_reassemble_the_frame(tmp_path, oracle, schedule, gather_mode, workers):
    """workers > 1: several batches in flight per rank (one host thread each) claim from the same counter."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), schedule, str(tmp_path), gather_mode, workers), nprocs=world, join=True)
    from raingun_b200.synth import make_scene

    data, _ = make_scene("C4", spheres=40, depth=4)
    ref, _, _ = oracle.render(data, W, H)
    for it in range(4):
        assert np.array_equal(np.load(tmp_path / f"frame{it}.npy"), ref), f"frame {it}"
    t0, t1 = np.load(tmp_path / "tiles0.npy"), np.load(tmp_path / "tiles1.npy")
    assert sorted(t0.tolist() + t1.tolist()) == list(range(n_tiles(H, TILE)))   # every tile exactly once
    if schedule == "static":
        assert sorted(t0.tolist()) == static_chunk(n_tiles(H, TILE), 2, 0)


def test_schedules_cover_every_tile_once():
    for nt in (1, 7, 135, 270):
        for world in (1, 2, 4, 8):
            chunks = guided_chunks(nt, world)
            assert [t for c in chunks for t in c] == list(range(nt))
            assert all(len(a) >= len(b) for a, b in zip(chunks, chunks[1:]))          # large first, small last
            assert sorted(t for r in range(world) for t in static_chunk(nt, world, r)) == list(range(nt))
            per_rank, tail = hybrid_plan(nt, world)
            covered = sorted([t for r in per_rank for t in r] + [t for c in tail for t in c])
            assert covered == list(range(nt)) and len(per_rank) == max(world, 1)
    from raingun_b200.dist import resolve_schedule
    assert resolve_schedule("auto", 270, 8) == "static" and resolve_schedule("auto", 270, 32) == "steal"
    assert resolve_schedule("steal", 270, 8) == "steal"
    assert rows_of_tiles([0, 13], 4, 54).tolist() == [0, 1, 2, 3, 52, 53]             # ragged last tile
    assert rows_of_tiles([], 4, 54).size == 0
