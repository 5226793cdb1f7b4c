"""ctypes wrapper around oracle/raingun_oracle.cpp (test infrastructure only)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

from raingun_b200.scene import SceneData, SceneDesc, Stats

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraingun_oracle.so")

RG_ERRORS = {0: "ok", -1: "invalid", -2: "portrait", -3: "too_large", -4: "depth"}


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "raingun_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "raingun_b200.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr) if os.path.exists(p))
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


class Oracle:
    def __init__(self) -> None:
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.rgo_render_rows.restype = ctypes.c_int
        L.rgo_render_rows.argtypes = [ctypes.POINTER(SceneDesc), ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(Stats)]
        L.rgo_fov_adjustment.restype = ctypes.c_double
        L.rgo_fov_adjustment.argtypes = [ctypes.c_double]
        L.rgo_texture_wrap.restype = ctypes.c_uint32
        L.rgo_texture_wrap.argtypes = [ctypes.c_float, ctypes.c_uint32]
        L.rgo_fresnel.restype = ctypes.c_double
        L.rgo_fresnel.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]
        L.rgo_quantise.restype = ctypes.c_uint8
        L.rgo_quantise.argtypes = [ctypes.c_float]
        L.rgo_intersect.restype = ctypes.c_int
        L.rgo_intersect.argtypes = [ctypes.c_uint8, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_double)]
        L.rgo_hardware_threads.restype = ctypes.c_int

    def hardware_threads(self) -> int:
        return max(1, int(self.lib.rgo_hardware_threads()))

    def render_rows(self, scene: SceneData, width: int, height: int, y0: int, y1: int,
                    threads: Optional[int] = None, max_depth: Optional[int] = None,
                    want_f32: bool = False) -> Tuple[np.ndarray, Stats, Optional[np.ndarray]]:
        desc, keep = scene.to_desc()
        rows = max(0, y1 - y0)
        out = np.zeros((rows, width, 4), np.uint8)
        f32 = np.zeros((rows, width, 3), np.float32) if want_f32 else None
        st = Stats()
        if threads is None:
            threads = self.hardware_threads()
        rc = self.lib.rgo_render_rows(ctypes.byref(desc), width, height, y0, y1,
                                      0xFFFFFFFF if max_depth is None else int(max_depth), int(threads),
                                      out.ctypes.data, f32.ctypes.data if want_f32 else None,
                                      ctypes.byref(st))
        del keep
        if rc != 0:
            raise RuntimeError(f"oracle render failed: {RG_ERRORS.get(rc, rc)}")
        return out, st, f32

    def render(self, scene: SceneData, width: int, height: int, **kw):
        return self.render_rows(scene, width, height, 0, height, **kw)

    def texture_wrap(self, val: float, size: int) -> int:
        return int(self.lib.rgo_texture_wrap(ctypes.c_float(val), size))

    def fresnel(self, incident, normal, index: float) -> float:
        i = np.ascontiguousarray(incident, np.float64)
        n = np.ascontiguousarray(normal, np.float64)
        return float(self.lib.rgo_fresnel(i.ctypes.data, n.ctypes.data, ctypes.c_float(index)))

    def quantise(self, c: float) -> int:
        return int(self.lib.rgo_quantise(ctypes.c_float(c)))

    def intersect(self, kind: int, geom8, origin, direction) -> Optional[float]:
        g = np.zeros(8, np.float64)
        g[: len(geom8)] = geom8
        o = np.ascontiguousarray(origin, np.float64)
        d = np.ascontiguousarray(direction, np.float64)
        t = ctypes.c_double(0.0)
        hit = self.lib.rgo_intersect(kind, g.ctypes.data, o.ctypes.data, d.ctypes.data, ctypes.byref(t))
        return float(t.value) if hit else None


_singleton: Optional[Oracle] = None


def oracle() -> Oracle:
    global _singleton
    if _singleton is None:
        _singleton = Oracle()
    return _singleton
