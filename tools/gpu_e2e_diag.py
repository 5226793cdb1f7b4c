import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import raingun_b200 as rg
from raingun_b200.synth import make_scene
from raingun_b200.examples import bundled_texture_loader
sd, spec = make_scene(sys.argv[1] if len(sys.argv) > 1 else "C4", texture_loader=bundled_texture_loader)
w, h = spec.width, spec.height
host = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
for it in range(6):
    t0 = time.perf_counter()
    desc, keep = sd.to_desc()
    t1 = time.perf_counter()
    sc = rg.Scene(sd)
    t2 = time.perf_counter()
    st = sc.render_rows_into(w, h, 0, h, host.data_ptr())
    t3 = time.perf_counter()
    sc.close()
    t4 = time.perf_counter()
    print(f"it{it} to_desc {1e3*(t1-t0):.2f} create {1e3*(t2-t1):.2f} render+d2h {1e3*(t3-t2):.2f} (dev {st.ms_device:.2f} wall {st.ms_wall:.2f}) close {1e3*(t4-t3):.2f} total {1e3*(t4-t0):.2f}", flush=True)
