"""torchrun check of the peer gather: every rank stores its rows into rank 0's frame over NVLink;
rank 0 compares the assembled frame with its own full render."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import raingun_b200 as rg
from raingun_b200.synth import make_scene
from raingun_b200.dist import PeerFrames, render_frame_sharded

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sd, spec = make_scene("C4", spheres=2000, depth=6)
w, h = 1920, 1080
sc = rg.Scene(sd, device=lr)
pf = PeerFrames(w, h, rank, world, lr)
ok = True
for frame_id, schedule in ((1, "static"), (2, "steal"), (3, "auto")):
    res = render_frame_sharded(lambda rows, fptr: sc.render_rowlist_scatter(w, h, rows, fptr, 0), w, h, rank, world,
                               frame_id, dev, schedule=schedule, gather_mode="peer", peer_frames=pf)
    if rank == 0:
        full = sc.render_image(w, h)
        same = np.array_equal(res.frame.cpu().numpy(), full)
        print(f"frame {frame_id} ({schedule}->{res.schedule}): assembled over {world} GPUs == single-GPU render: {same}", flush=True)
        ok = ok and same
    dist.barrier()
pf.close()
sc.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
