"""torchrun diagnostic: where a sharded frame's time goes (static batch / tail claims / gather)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import raingun_b200 as rg
from raingun_b200.synth import make_scene
from raingun_b200.dist import hybrid_plan, n_tiles, rows_of_tiles, TileCounter, default_store, _scatter_rows

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sd, spec = make_scene(sys.argv[1] if len(sys.argv) > 1 else "C4")
w, h = spec.width, spec.height
sc = rg.Scene(sd, device=lr)
staging = torch.empty(h * w * 4, dtype=torch.uint8, device=dev)
frame = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
nt = n_tiles(h, 8)
store = default_store()
sptr = torch.cuda.current_stream(dev).cuda_stream
for it in range(6):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    T = {}
    t0 = time.perf_counter()
    per_rank, tail = hybrid_plan(nt, world)
    counter = TileCounter(store, f"diag/{it}", tail)
    rows = rows_of_tiles(per_rank[rank], 8, h)
    t1 = time.perf_counter()
    st = sc.render_rowlist_device(w, h, rows, staging.data_ptr(), sptr)
    t2 = time.perf_counter()
    T["plan"] = t1 - t0; T["static_wall"] = t2 - t1; T["static_dev"] = st.ms_device / 1e3; T["static_trace"] = st.ms_trace / 1e3
    filled = rows.size
    allrows = [rows]
    nclaim = 0; tdev = 0.0
    while True:
        c0 = time.perf_counter()
        tiles = counter.claim()
        c1 = time.perf_counter()
        T["claim"] = T.get("claim", 0) + (c1 - c0)
        if tiles is None: break
        r2 = rows_of_tiles(tiles, 8, h)
        st2 = sc.render_rowlist_device(w, h, r2, staging[filled * w * 4:].data_ptr(), sptr)
        T["tail_wall"] = T.get("tail_wall", 0) + (time.perf_counter() - c1)
        tdev += st2.ms_device / 1e3
        filled += r2.size; allrows.append(r2); nclaim += 1
    T["tail_dev"] = tdev
    t3 = time.perf_counter()
    frame.zero_()
    _scatter_rows(frame, np.concatenate(allrows), staging[: filled * w * 4], w)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    dist.reduce(frame, dst=0, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    T["scatter"] = t4 - t3; T["reduce"] = t5 - t4; T["total"] = t5 - t0
    if it >= 2:
        print(f"it{it} rank{rank}/{world} rows {rows.size}+{filled - rows.size} claims {nclaim} " +
              " ".join(f"{k}={v * 1e3:.2f}" for k, v in T.items()), flush=True)
dist.destroy_process_group()
