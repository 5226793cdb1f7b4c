import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import raingun_b200 as rg
from raingun_b200.synth import make_scene
from raingun_b200.dist import rows_of_tiles, n_tiles
sd, spec = make_scene("C4")
w, h = spec.width, spec.height
sc = rg.Scene(sd)
out = torch.empty(h * w * 4, dtype=torch.uint8, device="cuda")
nt = n_tiles(h, 8)
for frac in (1, 2, 4, 8, 16, 32):
    rows = rows_of_tiles(list(range(0, nt, frac)), 8, h)
    best = 1e9; bw = 1e9
    for it in range(4):
        t0 = time.perf_counter()
        st = sc.render_rowlist_device(w, h, rows, out.data_ptr(), 0)
        bw = min(bw, (time.perf_counter() - t0) * 1e3); best = min(best, st.ms_device)
    print(f"1/{frac} of the frame: {rows.size} rows, device {best:.2f} ms, wall {bw:.2f} ms, trace {st.ms_trace:.2f} ms, ideal {27.9/frac:.2f} ms, launches {st.gpu_launches}", flush=True)
