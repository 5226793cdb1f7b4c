"""A/B of library builds (tools/build_variants.sh) on the grid path: device ms per frame of C3/C4/C5, with a
CRC of the frame (all builds must agree).  Usage: python tools/gpu_grid_ab.py [name:bps ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, zlib
sys.path.insert(0, ".")
import torch
import raingun_b200 as rg
from raingun_b200.synth import make_scene
from raingun_b200.examples import bundled_texture_loader
out = None
for wl in sys.argv[1:]:
    sd, spec = make_scene(wl, texture_loader=bundled_texture_loader)
    w, h = spec.width, spec.height
    if out is None or out.numel() < w * h * 4:
        out = torch.empty(h * w * 4, dtype=torch.uint8, device="cuda")
    sc = rg.Scene(sd)
    best = 1e9
    for it in range(5 if wl != "C5" else 3):
        st = sc.render_rows_device(w, h, 0, h, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        best = min(best, st.ms_device)
    crc = zlib.crc32(out[: w * h * 4].cpu().numpy().tobytes())
    print(f"  {wl}: {best:8.3f} ms  {st.rays / best / 1e3:7.0f} Mrays/s  exact {st.exact_tests}  crc {crc:08x}", flush=True)
    sc.close()
'''
variants = sys.argv[1:] or ["default:"]   # name:RG_GRID_BPS (blocks of 32 threads per SM; empty = the library default, 28):ENV=value,...
wls = os.environ.get("AB_WORKLOADS", "C3 C4").split()
for v in variants:
    name, _, rest = v.partition(":")
    bps, _, extra = rest.partition(":")
    env = dict(os.environ)
    if name != "default":
        env["RAINGUN_B200_LIB"] = os.path.join(ROOT, "raingun_b200", "_variants", name + ".so")
    if bps:
        env["RG_GRID_BPS"] = bps
    for kv in extra.split(",") if extra else []:
        k, _, val = kv.partition("=")
        env[k] = val
    print(f"== {v}", flush=True)
    subprocess.run([sys.executable, "-c", CHILD] + wls, env=env, cwd=ROOT)
