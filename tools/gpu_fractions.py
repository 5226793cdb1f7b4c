"""Device ms of 1/N of the C4 frame as one interleaved row-list batch (the multi-GPU unit)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import raingun_b200 as rg
from raingun_b200.synth import make_scene
from raingun_b200.dist import rows_of_tiles, n_tiles
sd, spec = make_scene("C4")
w, h = spec.width, spec.height
out = torch.empty(h * w * 4, dtype=torch.uint8, device="cuda")
nt = n_tiles(h, 8)
sc = rg.Scene(sd)
full = None
for frac in (1, 2, 4, 8, 16):
    rows = rows_of_tiles(list(range(0, nt, frac)), 8, h)
    best = 1e9
    for _ in range(7):
        st = sc.render_rowlist_device(w, h, rows, out.data_ptr(), 0)
        best = min(best, st.ms_device)
    full = full or best
    print(f"1/{frac:<2d}: {best:7.3f} ms  (ideal {full / frac:6.3f}, efficiency {full / frac / best:.3f})", flush=True)
