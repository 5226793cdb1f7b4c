"""ctypes binding of libraingun_host.so — the native host side (include/raingun_host.h):
YAML scene loading, JPEG/PNG decode, PNG encode and the CLI option mapping, in C++.

The Python ingestion in scene.py stays as the second, independent implementation of the same
schema (tests compare the two field by field); textures are always decoded natively so that the
Python host, the C++ CLI and the oracle all see the same texels.
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, List, Optional, Tuple

import numpy as np

from .scene import SceneData, SceneDesc, SceneError, TextureDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libraingun_host.so")
CLI_PATH = os.path.join(_HERE, "host", "_build", "raingun")

OK, E_INVALID, E_IO, E_FORMAT, E_UNSUPPORTED, E_SCHEMA, E_USAGE = 0, -1, -2, -3, -4, -5, -6
ERRORS = {0: "RGH_OK", -1: "RGH_E_INVALID", -2: "RGH_E_IO", -3: "RGH_E_FORMAT", -4: "RGH_E_UNSUPPORTED",
          -5: "RGH_E_SCHEMA", -6: "RGH_E_USAGE"}

# every symbol include/raingun_host.h declares
EXPORTS = ("rgh_jpeg_decode", "rgh_png_decode", "rgh_bmp_decode", "rgh_tga_decode", "rgh_pnm_decode", "rgh_gif_decode",
           "rgh_png_encode", "rgh_image_open", "rgh_png_save", "rgh_free",
           "rgh_alloc", "rgh_scene_parse", "rgh_scene_load", "rgh_scene_desc", "rgh_scene_limit_depth",
           "rgh_scene_texture_path", "rgh_scene_destroy", "rgh_cli_parse", "rgh_last_error")


class Image(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("channels", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32), ("pixels", ctypes.c_void_p)]


class CliOptions(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("max_depth_limit", ctypes.c_int32),
                ("preview", ctypes.c_int32), ("input", ctypes.c_char * 4096), ("output", ctypes.c_char * 4096)]


TEXTURE_CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(Image), ctypes.c_void_p)


class HostError(SceneError):
    def __init__(self, code: int, message: str) -> None:
        super().__init__(message)
        self.code = code


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C raingun_b200/host lib` "
                          "(or python -c 'import __graft_entry__ as g; g.build()')")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, sz = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t
    L.rgh_jpeg_decode.restype = ctypes.c_int
    L.rgh_jpeg_decode.argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(Image)]
    L.rgh_png_decode.restype = ctypes.c_int
    L.rgh_png_decode.argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(Image)]
    for name in ("rgh_bmp_decode", "rgh_tga_decode", "rgh_pnm_decode", "rgh_gif_decode"):
        getattr(L, name).restype = ctypes.c_int
        getattr(L, name).argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(Image)]
    L.rgh_png_encode.restype = ctypes.c_int
    L.rgh_png_encode.argtypes = [vp, u32, u32, u32, ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.rgh_image_open.restype = ctypes.c_int
    L.rgh_image_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(Image)]
    L.rgh_png_save.restype = ctypes.c_int
    L.rgh_png_save.argtypes = [ctypes.c_char_p, vp, u32, u32, u32]
    L.rgh_free.restype = None
    L.rgh_free.argtypes = [vp]
    L.rgh_alloc.restype = vp
    L.rgh_alloc.argtypes = [sz]
    L.rgh_scene_parse.restype = ctypes.c_int
    L.rgh_scene_parse.argtypes = [ctypes.c_char_p, sz, ctypes.c_char_p, TEXTURE_CB, vp, ctypes.POINTER(vp)]
    L.rgh_scene_load.restype = ctypes.c_int
    L.rgh_scene_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(vp)]
    L.rgh_scene_desc.restype = ctypes.POINTER(SceneDesc)
    L.rgh_scene_desc.argtypes = [vp]
    L.rgh_scene_limit_depth.restype = None
    L.rgh_scene_limit_depth.argtypes = [vp, u32]
    L.rgh_scene_texture_path.restype = ctypes.c_char_p
    L.rgh_scene_texture_path.argtypes = [vp, u32]
    L.rgh_scene_destroy.restype = None
    L.rgh_scene_destroy.argtypes = [vp]
    L.rgh_cli_parse.restype = ctypes.c_int
    L.rgh_cli_parse.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(CliOptions)]
    L.rgh_last_error.restype = ctypes.c_char_p
    L.rgh_last_error.argtypes = []
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != OK:
        msg = lib().rgh_last_error()
        raise HostError(rc, msg.decode("utf-8", "replace") if msg else ERRORS.get(rc, str(rc)))


def _take_image(img: Image) -> np.ndarray:
    n = img.width * img.height * img.channels
    buf = (ctypes.c_uint8 * n).from_address(img.pixels)
    out = np.frombuffer(buf, np.uint8).reshape(img.height, img.width, img.channels).copy()
    lib().rgh_free(img.pixels)
    return out


def decode_jpeg(data: bytes) -> np.ndarray:
    """(H, W, 1|3) uint8, decoded with jpeg-decoder 0.1.11's arithmetic (rgh_jpeg.cpp)."""
    img = Image()
    _check(lib().rgh_jpeg_decode(data, len(data), ctypes.byref(img)))
    return _take_image(img)


def decode_png(data: bytes) -> np.ndarray:
    img = Image()
    _check(lib().rgh_png_decode(data, len(data), ctypes.byref(img)))
    return _take_image(img)


def _decode_with(name: str, data: bytes) -> np.ndarray:
    img = Image()
    _check(getattr(lib(), name)(data, len(data), ctypes.byref(img)))
    return _take_image(img)


def decode_bmp(data: bytes) -> np.ndarray:
    return _decode_with("rgh_bmp_decode", data)


def decode_tga(data: bytes) -> np.ndarray:
    return _decode_with("rgh_tga_decode", data)


def decode_pnm(data: bytes) -> np.ndarray:
    return _decode_with("rgh_pnm_decode", data)


def decode_gif(data: bytes) -> np.ndarray:
    return _decode_with("rgh_gif_decode", data)


def decode_image(data: bytes) -> np.ndarray:
    """Texture bytes -> (H, W, 3|4) uint8 as `DynamicImage::get_pixel` presents them."""
    if data[:2] == b"\xff\xd8":
        a = decode_jpeg(data)
    elif data[:8] == b"\x89PNG\r\n\x1a\n":
        a = decode_png(data)
    elif data[:2] == b"BM":
        a = decode_bmp(data)
    elif data[:6] in (b"GIF87a", b"GIF89a"):
        a = decode_gif(data)
    elif data[:1] == b"P" and data[1:2] in b"123456":
        a = decode_pnm(data)
    else:
        raise HostError(E_UNSUPPORTED, "unsupported image format (jpg and png are built)")
    return np.repeat(a, 3, axis=2) if a.shape[2] == 1 else a


def open_image(path: str) -> np.ndarray:
    """image::open (material.rs:42): decoder chosen by extension."""
    img = Image()
    _check(lib().rgh_image_open(os.fsencode(path), ctypes.byref(img)))
    a = _take_image(img)
    return np.repeat(a, 3, axis=2) if a.shape[2] == 1 else a


def encode_png(pixels: np.ndarray) -> bytes:
    a = np.ascontiguousarray(pixels, np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    out, n = ctypes.c_void_p(), ctypes.c_size_t()
    _check(lib().rgh_png_encode(a.ctypes.data, a.shape[1], a.shape[0], a.shape[2], ctypes.byref(out), ctypes.byref(n)))
    data = ctypes.string_at(out.value, n.value)
    lib().rgh_free(out)
    return data


def save_png(path: str, pixels: np.ndarray) -> None:
    a = np.ascontiguousarray(pixels, np.uint8)
    _check(lib().rgh_png_save(os.fsencode(path), a.ctypes.data, a.shape[1], a.shape[0], a.shape[2]))


def _desc_to_scene_data(d: SceneDesc, paths: List[str]) -> SceneData:
    def arr(ptr, dtype, shape):
        n = int(np.prod(shape))
        if n == 0 or not ptr:
            return np.zeros(shape, dtype)
        nbytes = n * np.dtype(dtype).itemsize
        return np.frombuffer(ctypes.string_at(ptr, nbytes), dtype).reshape(shape).copy()

    n, nl = d.n_bodies, d.n_lights
    sd = SceneData()
    sd.fov = d.fov
    sd.default_color = np.array(list(d.default_color), np.float32)
    sd.max_recursion_depth = d.max_recursion_depth
    sd.body_kind = arr(d.body_kind, np.uint8, (n,))
    sd.body_geom = arr(d.body_geom, np.float64, (n, 8))
    sd.coloration_kind = arr(d.coloration_kind, np.uint8, (n,))
    sd.color = arr(d.color, np.float32, (n, 3))
    sd.texture_id = arr(d.texture_id, np.int32, (n,))
    sd.texture_offset = arr(d.texture_offset, np.float32, (n, 2))
    sd.albedo = arr(d.albedo, np.float32, (n,))
    sd.surface_kind = arr(d.surface_kind, np.uint8, (n,))
    sd.surface_param = arr(d.surface_param, np.float32, (n, 2))
    sd.light_kind = arr(d.light_kind, np.uint8, (nl,))
    sd.light_vec = arr(d.light_vec, np.float64, (nl, 3))
    sd.light_color = arr(d.light_color, np.float32, (nl, 3))
    sd.light_intensity = arr(d.light_intensity, np.float32, (nl,))
    tds = ctypes.cast(d.textures, ctypes.POINTER(TextureDesc)) if d.n_textures else None
    for i in range(d.n_textures):
        t = tds[i]
        sd.textures.append(arr(t.pixels, np.uint8, (t.height, t.width, t.channels)))
    sd.texture_names = list(paths)
    return sd


def parse_scene(yaml_text: str, texture_loader: Optional[Callable[[str], np.ndarray]] = None,
                texture_root: Optional[str] = None, max_depth_limit: Optional[int] = None) -> SceneData:
    """serde_yaml::from_reader::<Scene> (src/main.rs:118) done by the C++ host library; returns the
    same SceneData the Python ingestion produces. `texture_loader(path) -> (H, W, 3|4) uint8` overrides
    the native image_open (used for the bundled example textures)."""
    L = lib()
    raw = yaml_text.encode("utf-8") if isinstance(yaml_text, str) else bytes(yaml_text)
    failure: List[BaseException] = []

    def _cb(path, out, _user):
        try:
            a = np.ascontiguousarray(texture_loader(path.decode("utf-8")), np.uint8)
            if a.ndim != 3 or a.shape[2] not in (1, 3, 4):
                raise ValueError("texture loader must return (H, W, 1|3|4) uint8")
            p = L.rgh_alloc(a.size)
            ctypes.memmove(p, a.ctypes.data, a.size)
            out[0].height, out[0].width, out[0].channels = a.shape
            out[0].pixels = p
            return 0
        except BaseException as e:  # noqa: BLE001 - reported through the C error path
            failure.append(e)
            return 1

    cb = TEXTURE_CB(_cb) if texture_loader is not None else ctypes.cast(None, TEXTURE_CB)
    handle = ctypes.c_void_p()
    rc = L.rgh_scene_parse(raw, len(raw), os.fsencode(texture_root) if texture_root else None, cb, None,
                           ctypes.byref(handle))
    if rc != OK:
        msg = L.rgh_last_error().decode("utf-8", "replace")
        if failure:
            msg += f" ({failure[0]})"
        raise HostError(rc, msg)
    try:
        if max_depth_limit is not None:
            L.rgh_scene_limit_depth(handle, int(max_depth_limit))
        d = L.rgh_scene_desc(handle)[0]
        paths = []
        for i in range(d.n_textures):
            paths.append(L.rgh_scene_texture_path(handle, i).decode("utf-8"))
        return _desc_to_scene_data(d, paths)
    finally:
        L.rgh_scene_destroy(handle)


def load_scene(path: str, texture_root: Optional[str] = None) -> SceneData:
    with open(path, "rb") as f:
        return parse_scene(f.read().decode("utf-8"), None, texture_root)


def cli_parse(argv: List[str]) -> Tuple[int, int, Optional[int], bool, str, str]:
    """RenderOptions::from + output-path defaulting (src/main.rs:66-113).
    Returns (width, height, max_depth_limit | None, preview, input, output)."""
    arr = (ctypes.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    opt = CliOptions()
    _check(lib().rgh_cli_parse(len(argv), arr, ctypes.byref(opt)))
    return (opt.width, opt.height, None if opt.max_depth_limit < 0 else opt.max_depth_limit, bool(opt.preview),
            os.fsdecode(opt.input), os.fsdecode(opt.output))
