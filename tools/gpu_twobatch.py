"""Does a second in-flight batch hide the latency-bound tails of the first?  The share one GPU would own
at N GPUs (every N-th 8-row tile of the C4 frame) rendered as ONE batch vs as K concurrent batches
(K scene handles, K streams, K host threads), wall clock of the whole share."""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np, torch
import raingun_b200 as rg
from raingun_b200 import _native as N
from raingun_b200.synth import make_scene
from raingun_b200.dist import rows_of_tiles, n_tiles

sd, spec = make_scene("C4")
w, h = spec.width, spec.height
nt = n_tiles(h, 8)
K = 3
scenes = [rg.Scene(sd) for _ in range(K)]
streams = [torch.cuda.Stream() for _ in range(K)]
outs = [torch.empty(h * w * 4, dtype=torch.uint8, device="cuda") for _ in range(K)]


def run(parts):
    def work(k):
        scenes[k].render_rowlist_device(w, h, parts[k], outs[k].data_ptr(), streams[k].cuda_stream)
    th = [threading.Thread(target=work, args=(k,)) for k in range(len(parts))]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


for frac in (1, 2, 4, 8, 16):
    tiles = list(range(0, nt, frac))
    for k in (1, 2, 3):
        parts = [rows_of_tiles(tiles[i::k], 8, h) for i in range(k)]
        best = min(run(parts) for _ in range(6))
        print(f"1/{frac} of C4 as {k} concurrent batch(es): wall {best:.3f} ms (ideal {26.75 / frac:.2f})", flush=True)
