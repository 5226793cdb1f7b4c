import sys
sys.path.insert(0, ".")
import numpy as np
import raingun_b200 as rg
from raingun_b200.synth import make_scene
d, _ = make_scene("C4", spheres=400, depth=3)
sc = rg.Scene(d); sc.set_accel(1); sc.set_option(5, 2)
img = sc.render_image(256, 144); st = sc.last_stats
print("mismatches", st.cull_unsound, "rays", st.rays_primary, st.rays_shadow, st.rays_reflection, st.rays_transmission)
