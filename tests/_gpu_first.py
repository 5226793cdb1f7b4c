import sys, time
sys.path.insert(0, ".")
import numpy as np
import raingun_b200 as rg
from raingun_b200.examples import example_scene, example_golden
from oracle import oracle
O = oracle()
print("devices", rg.device_count())
print("peaks", rg.measure_peaks(0))
for name in ("test1", "test2", "test3"):
    sd = example_scene(name)
    ref, ost, _ = O.render(sd, 800, 600)
    sc = rg.Scene(sd)
    sc.set_pipeline(rg.PIPELINE_MEGAKERNEL)
    img = sc.render_image(800, 600)
    st = sc.last_stats
    img = sc.render_image(800, 600)
    st = sc.last_stats
    diff = np.abs(img.astype(int) - ref.astype(int)).max(axis=2)
    print(name, "mega vs oracle: differing px", int((diff > 0).sum()), "max", diff.max(), "ms", st.ms_device,
          "rays", st.rays_primary, st.rays_shadow, st.rays_reflection, st.rays_transmission,
          "oracle", ost.rays_primary, ost.rays_shadow, ost.rays_reflection, ost.rays_transmission)
