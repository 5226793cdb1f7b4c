"""RG_OPT_VERIFY_CULL=2 on real sizes: every ray the grid tracer traced is re-traced with the verbatim f64 scan
on the device and compared (must print unsound 0)."""
import sys
sys.path.insert(0, ".")
import raingun_b200 as rg
from raingun_b200 import _native as N
from raingun_b200.synth import make_scene
from raingun_b200.examples import bundled_texture_loader
for wl, w, h, kw in (("C4", 960, 540, {}), ("C3", 960, 540, {}), ("C5", 480, 270, {})):
    sd, spec = make_scene(wl, texture_loader=bundled_texture_loader, **kw)
    with rg.Scene(sd) as sc:
        sc.set_accel(rg.ACCEL_GRID)
        sc.set_option(N.OPT_VERIFY_CULL, 2)
        sc.render_image(w, h)
        st = sc.last_stats
        print(f"{wl} {w}x{h} bodies {sd.n_bodies}: rays {st.rays} verify mismatches {st.cull_unsound} device {st.ms_device:.1f} ms", flush=True)
