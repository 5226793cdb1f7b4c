import sys
sys.path.insert(0, ".")
import numpy as np
import raingun_b200 as rg
from raingun_b200.examples import example_scene, bundled_texture_loader
from raingun_b200.synth import make_scene
from oracle import oracle
O = oracle()
cases = [(n, example_scene(n), 800, 600) for n in ("test1", "test2", "test3")]
d, _ = make_scene("C3", spheres=300, depth=4); cases.append(("C3s", d, 320, 180))
d, _ = make_scene("C4", spheres=400, depth=8); cases.append(("C4s", d, 256, 144))
for name, sd, w, h in cases:
    ref, ost, _ = O.render(sd, w, h)
    for label, pipe, acc in (("mega", 1, 1), ("wf-brute", 0, 1), ("wf-grid", 0, 2)):
        sc = rg.Scene(sd)
        sc.set_pipeline(pipe); sc.set_accel(acc)
        sc.set_option(5, 2 if pipe == 0 else 0)
        img = sc.render_image(w, h)
        st = sc.last_stats
        diff = np.abs(img.astype(int) - ref.astype(int)).max(axis=2)
        ys, xs = np.nonzero(diff)
        print(name, label, "diff px", int((diff > 0).sum()), "max", diff.max(),
              "rays", st.rays_primary, st.rays_shadow, st.rays_reflection, st.rays_transmission,
              "oracle", ost.rays_primary, ost.rays_shadow, ost.rays_reflection, ost.rays_transmission,
              "unsound", st.cull_unsound, "exact", st.exact_tests, "accel", st.accel_used, "lvl", st.max_level,
              "first diff", (ys[:3].tolist(), xs[:3].tolist()) if len(ys) else None, flush=True)
        sc.close()
