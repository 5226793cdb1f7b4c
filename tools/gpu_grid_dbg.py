"""Phase / exact-test breakdown of k_trace_grid (library built with -DRG_GRID_DEBUG=1 by tools/build_variants.sh dbg ...;
run with RAINGUN_B200_LIB=raingun_b200/_variants/dbg.so).  One instrumented frame (RG_OPT_TRACE_STATS) per workload."""
import ctypes, sys
sys.path.insert(0, ".")
import torch
import raingun_b200 as rg
from raingun_b200 import _native as N
from raingun_b200.synth import make_scene
from raingun_b200.examples import bundled_texture_loader
lib = N.lib()
names = ["cyc_refill", "cyc_scan", "cyc_exact", "exact_phases", "waiting_lanes_sum", "exact_tests", "x_cull_false_pos", "x_behind",
         "x_hit_not_better", "x_hit_useful", "x_duplicate", "x_behind_on_surface", "refill_idle_sum", "loop_iters", "", ""]
for wl in sys.argv[1:] or ["C4"]:
    sd, spec = make_scene(wl, texture_loader=bundled_texture_loader)
    w, h = spec.width, spec.height
    out = torch.empty(h * w * 4, dtype=torch.uint8, device="cuda")
    sc = rg.Scene(sd)
    sc.set_option(N.OPT_TRACE_STATS, 1)
    buf = (ctypes.c_ulonglong * 32)()
    have_dbg = hasattr(lib, "rg_debug_grid_counters")
    if have_dbg: lib.rg_debug_grid_counters(buf, 1)
    st = sc.render_rows_device(w, h, 0, h, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if have_dbg: lib.rg_debug_grid_counters(buf, 1)
    print(wl, "rays", st.rays, "ms", st.ms_device, "exact", st.exact_tests, "| per ray: cells %.2f records %.2f culls %.2f; rays per refill %.1f; scan lane use %.3f" % (
        st.grid_cells / st.rays, st.grid_fetches / st.rays, st.grid_culls / st.rays, st.rays / max(1, st.grid_refills), st.grid_lane_steps / max(1, st.grid_lane_slots)))
    if not have_dbg:
        sc.close()
        continue
    for v, label in ((0, "nearest"), (1, "any-hit")):
        d = {names[k]: buf[v * 16 + k] for k in range(14)}
        cyc = d["cyc_refill"] + d["cyc_scan"] + d["cyc_exact"]
        print(f"  {label}: phase cycles refill {d['cyc_refill']/cyc:.3f} scan {d['cyc_scan']/cyc:.3f} exact {d['cyc_exact']/cyc:.3f};"
              f" lanes per exact phase {d['waiting_lanes_sum']/max(1,d['exact_phases']):.2f}; exact phases per loop iter {d['exact_phases']/max(1,d['loop_iters']):.3f};"
              f" idle lanes per refill {d['refill_idle_sum']/max(1,st.grid_refills):.2f}")
        n = max(1, d["exact_tests"])
        print(f"    exact tests {d['exact_tests']}: cull false positive {d['x_cull_false_pos']/n:.3f}, behind {d['x_behind']/n:.3f} (origin on the surface {d['x_behind_on_surface']/n:.3f}),"
              f" hit not better {d['x_hit_not_better']/n:.3f}, useful {d['x_hit_useful']/n:.3f}; duplicates {d['x_duplicate']/n:.3f}")
    sc.close()
