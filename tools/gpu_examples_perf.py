import sys, time
sys.path.insert(0, ".")
import torch
import raingun_b200 as rg
from raingun_b200.examples import example_scene
out = torch.empty(3840 * 2160 * 4, dtype=torch.uint8, device="cuda")
for name in ("test1", "test2", "test3"):
    sd = example_scene(name)
    for label, pipe in (("wavefront", rg.PIPELINE_WAVEFRONT), ("megakernel", rg.PIPELINE_MEGAKERNEL)):
        for (w, h) in ((800, 600), (1920, 1080), (3840, 2160)):
            with rg.Scene(sd) as sc:
                sc.set_pipeline(pipe)
                best = 1e9
                for it in range(6):
                    st = sc.render_rows_device(w, h, 0, h, out.data_ptr(), 0)
                    best = min(best, st.ms_device)
                print(f"{name} {w}x{h} {label}: dev {best:.3f} ms rays {st.rays} -> {st.rays/best/1e3:.0f} Mrays/s, launches {st.gpu_launches}, levels {st.max_level+1}, replays {st.graph_replays}", flush=True)
