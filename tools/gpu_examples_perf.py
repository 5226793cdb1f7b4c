import sys, time
sys.path.insert(0, ".")
import raingun_b200 as rg
from raingun_b200.examples import example_scene
for name in ("test1", "test2", "test3"):
    sd = example_scene(name)
    for (w, h) in ((800, 600), (3840, 2160)):
        with rg.Scene(sd) as sc:
            for it in range(4):
                t0 = time.perf_counter(); sc.render_image(w, h); wall = time.perf_counter() - t0
            st = sc.last_stats
            print(f"{name} {w}x{h}: dev {st.ms_device:.3f} ms wall {wall*1e3:.2f} ms rays {st.rays} -> {st.rays/st.ms_device/1e3:.0f} Mrays/s (device), launches {st.gpu_launches}, levels {st.max_level+1}, accel {st.accel_used}", flush=True)
