import sys, time
sys.path.insert(0, ".")
import numpy as np
import raingun_b200 as rg
from raingun_b200.synth import make_scene
print("peaks", rg.measure_peaks(0))
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C3", "C4"]
for name in which:
    t0 = time.time()
    from raingun_b200.examples import bundled_texture_loader
    sd, spec = make_scene(name, texture_loader=bundled_texture_loader)
    print(name, "scene built in %.1fs" % (time.time() - t0), "bodies", sd.n_bodies, flush=True)
    w, h = spec.width, spec.height
    imgs = {}
    for label, pipe, acc in ((("wf-grid", 0, 2),) if name == "C5" and "--brute" not in sys.argv else (("wf-grid", 0, 2), ("wf-brute", 0, 1))):
        sc = rg.Scene(sd)
        sc.set_pipeline(pipe); sc.set_accel(acc)
        for it in range(3):
            t0 = time.time()
            img = sc.render_image(w, h)
            wall = time.time() - t0
            st = sc.last_stats
            print(name, label, "iter", it, "wall %.1f ms" % (wall * 1e3), "dev %.2f ms trace %.2f ms" % (st.ms_device, st.ms_trace),
                  "rays", st.rays, "(p %d s %d r %d t %d)" % (st.rays_primary, st.rays_shadow, st.rays_reflection, st.rays_transmission),
                  "Mrays/s %.1f" % (st.rays / st.ms_device / 1e3), "launches", st.gpu_launches, "lvl", st.max_level,
                  "exact", st.exact_tests, "tests/s %.3g" % (st.body_tests / (st.ms_device * 1e-3)), flush=True)
        imgs[label] = img
        sc.close()
    if "wf-brute" not in imgs: continue
    a, b = imgs["wf-grid"], imgs["wf-brute"]
    print(name, "grid vs brute differing px:", int((np.abs(a.astype(int) - b.astype(int)).max(axis=2) > 0).sum()), flush=True)
