"""In-library multi-GPU: wall clock of rg_render on a multi-device scene vs the devices' own event times."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import raingun_b200 as rg
from raingun_b200 import _native as N
from raingun_b200.synth import make_scene
sd, spec = make_scene("C4")
w, h = spec.width, spec.height
nd = rg.device_count()
host = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
pageable = np.empty((h, w, 4), np.uint8)
for devices in ([0], list(range(nd)), [0, 0]):
    for schedule in (1, 2):
        with rg.Scene(sd, devices=devices) as sc:
            sc.set_option(N.OPT_SCHEDULE, schedule)
            for target, ptr in (("pinned", host.data_ptr()), ("pageable", pageable.ctypes.data)):
                for it in range(5):
                    t0 = time.perf_counter()
                    st = sc.render_rows_into(w, h, 0, h, ptr)
                    wall = (time.perf_counter() - t0) * 1e3
                print(f"devices {devices} schedule {schedule} -> {target}: wall {wall:.2f} ms, lib wall {st.ms_wall:.2f}, max device {st.ms_device:.2f} ms, "
                      f"batches {st.batches}, used {st.devices_used}, replays {st.graph_replays}", flush=True)
# one device rendering half the frame into host memory through the row-list entry point
rows = np.concatenate([np.arange(y, min(h, y + 8), dtype=np.uint32) for y in range(0, h, 16)])
with rg.Scene(sd, device=0) as sc:
    for it in range(5):
        t0 = time.perf_counter()
        st = sc.render_rowlist_host(w, h, rows, host.data_ptr())
        wall = (time.perf_counter() - t0) * 1e3
    print(f"one device, every other tile -> pinned host frame: wall {wall:.2f} ms, device {st.ms_device:.2f} ms", flush=True)
