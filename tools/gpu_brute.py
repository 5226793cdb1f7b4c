import sys, zlib
sys.path.insert(0, ".")
import numpy as np
import raingun_b200 as rg
from raingun_b200.synth import make_scene
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
from raingun_b200.examples import bundled_texture_loader
sd, spec = make_scene(name, texture_loader=bundled_texture_loader)
w, h = spec.width, spec.height
with rg.Scene(sd) as sc:
    sc.set_accel(rg.ACCEL_BRUTE)
    for it in range(3):
        img = sc.render_image(w, h)
        st = sc.last_stats
print(name, "brute dev %.2f ms trace %.2f ms rays %d exact %d crc %08x" % (st.ms_device, st.ms_trace, st.rays, st.exact_tests, zlib.crc32(img.tobytes())), flush=True)
