"""Where does the per-pixel megakernel stop beating the wavefront?  C4-style scenes of N spheres at 1080p / 4K."""
import sys
sys.path.insert(0, ".")
import torch
import raingun_b200 as rg
from raingun_b200.synth import make_scene
out = torch.empty(3840 * 2160 * 4, dtype=torch.uint8, device="cuda")
for n in (4, 8, 12, 16, 24, 32, 48, 64, 96):
    sd, _ = make_scene("C4", spheres=n, depth=8)
    for (w, h) in ((1920, 1080), (3840, 2160)):
        res = {}
        for label, pipe in (("wavefront", rg.PIPELINE_WAVEFRONT), ("megakernel", rg.PIPELINE_MEGAKERNEL)):
            with rg.Scene(sd) as sc:
                sc.set_pipeline(pipe)
                best = 1e9
                for it in range(5):
                    st = sc.render_rows_device(w, h, 0, h, out.data_ptr(), 0)
                    best = min(best, st.ms_device)
                res[label] = best
        print(f"{n:3d} spheres {w}x{h}: wavefront {res['wavefront']:.3f} ms  megakernel {res['megakernel']:.3f} ms  rays {st.rays}", flush=True)
