"""Scene ingestion: the reference's YAML schema -> flattened, body-indexed arrays.

This is the host-side "scene-upload layer": it accepts exactly what the
reference's serde derive accepts and produces the arrays of ``rg_scene_desc``
(include/raingun_b200.h).

Reference anchors (all under /root/reference):
  * root struct, camelCase keys, deny_unknown_fields, defaults fov=90 / depth=10 /
    black background ........................... raingun-lib/src/scene.rs:11-31
  * externally tagged enums Body / Light / Coloration / Surface
    ............................................ bodies.rs:41-47, lights.rs:22-26,
                                                 material.rs:20-24,49-54
  * vectors as ``{x,y,z}`` maps or ``[x,y,z]`` sequences (cgmath "eders")
    ............................................ examples/test1.yml:5-8 vs :13
  * unit variants written ``Diffuse`` or ``Diffuse:`` .. examples/test1.yml:36 vs :44-45
  * colours ``"#rrggbb"`` only, byte/255 in f32 ........ color.rs:114-130
  * texture keys image / x_offset / y_offset, path relative to the CWD
    ............................................ material.rs:26-47
  * f32 fields arrive as f64 and are narrowed .......... material.rs:10,30-31,52-53;
                                                         lights.rs:12,19
  * ``--draft`` style depth override ................... src/main.rs:119-123
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

BODY_SPHERE, BODY_PLANE, BODY_DISK, BODY_AABB = 0, 1, 2, 3
COLORATION_COLOR, COLORATION_TEXTURE = 0, 1
SURFACE_DIFFUSE, SURFACE_REFLECTING, SURFACE_REFRACTIVE = 0, 1, 2
LIGHT_DIRECTIONAL, LIGHT_SPHERICAL = 0, 1
ABI_VERSION = 2


class SceneError(ValueError):
    """The YAML does not deserialize into the reference's ``Scene``."""


# --------------------------------------------------------------------------- ctypes
class TextureDesc(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_uint32),
        ("height", ctypes.c_uint32),
        ("channels", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
        ("pixels", ctypes.c_void_p),
    ]


class SceneDesc(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_uint32),
        ("max_recursion_depth", ctypes.c_uint32),
        ("fov", ctypes.c_double),
        ("default_color", ctypes.c_float * 3),
        ("n_bodies", ctypes.c_uint32),
        ("body_kind", ctypes.c_void_p),
        ("body_geom", ctypes.c_void_p),
        ("coloration_kind", ctypes.c_void_p),
        ("color", ctypes.c_void_p),
        ("texture_id", ctypes.c_void_p),
        ("texture_offset", ctypes.c_void_p),
        ("albedo", ctypes.c_void_p),
        ("surface_kind", ctypes.c_void_p),
        ("surface_param", ctypes.c_void_p),
        ("n_lights", ctypes.c_uint32),
        ("n_textures", ctypes.c_uint32),
        ("light_kind", ctypes.c_void_p),
        ("light_vec", ctypes.c_void_p),
        ("light_color", ctypes.c_void_p),
        ("light_intensity", ctypes.c_void_p),
        ("textures", ctypes.c_void_p),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("rays_primary", ctypes.c_uint64),
        ("rays_shadow", ctypes.c_uint64),
        ("rays_reflection", ctypes.c_uint64),
        ("rays_transmission", ctypes.c_uint64),
        ("body_tests", ctypes.c_uint64),
        ("exact_tests", ctypes.c_uint64),
        ("cull_unsound", ctypes.c_uint64),
        ("err_nan_distance", ctypes.c_uint64),
        ("err_transmission_none", ctypes.c_uint64),
        ("err_aabb_normal", ctypes.c_uint64),
        ("ms_device", ctypes.c_double),
        ("ms_trace", ctypes.c_double),
        ("ms_wall", ctypes.c_double),
        ("gpu_launches", ctypes.c_uint32),
        ("batches", ctypes.c_uint32),
        ("max_level", ctypes.c_uint32),
        ("accel_used", ctypes.c_uint32),
        ("host_free", ctypes.c_uint32),
        ("graph_replays", ctypes.c_uint32),
        ("grid_cells", ctypes.c_uint64),
        ("grid_fetches", ctypes.c_uint64),
        ("grid_culls", ctypes.c_uint64),
        ("grid_refills", ctypes.c_uint64),
        ("grid_lane_steps", ctypes.c_uint64),
        ("grid_lane_slots", ctypes.c_uint64),
        ("pipeline_used", ctypes.c_uint32),
        ("devices_used", ctypes.c_uint32),
    ]

    @property
    def rays(self) -> int:
        return self.rays_primary + self.rays_shadow + self.rays_reflection + self.rays_transmission

    def as_dict(self) -> Dict[str, Any]:
        return {name: getattr(self, name) for name, _ in self._fields_}


# --------------------------------------------------------------------------- data
@dataclass
class SceneData:
    """The reference's ``Scene`` (scene.rs:11-19), flattened body-by-body."""

    fov: float = 90.0
    default_color: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    max_recursion_depth: int = 10
    body_kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    body_geom: np.ndarray = field(default_factory=lambda: np.zeros((0, 8), np.float64))
    coloration_kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    color: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    texture_id: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    texture_offset: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))
    albedo: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    surface_kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    surface_param: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))
    light_kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    light_vec: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float64))
    light_color: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    light_intensity: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    textures: List[np.ndarray] = field(default_factory=list)  # (H, W, 3|4) uint8
    texture_names: List[str] = field(default_factory=list)

    @property
    def n_bodies(self) -> int:
        return int(self.body_kind.shape[0])

    @property
    def n_lights(self) -> int:
        return int(self.light_kind.shape[0])

    def with_max_depth_limit(self, limit: Optional[int]) -> "SceneData":
        """src/main.rs:119-123: a limit only ever lowers the scene's own depth."""
        import copy

        out = copy.copy(self)
        if limit is not None and limit < self.max_recursion_depth:
            out.max_recursion_depth = int(limit)
        return out

    def to_desc(self):
        """Returns ``(SceneDesc, keepalive)``; the desc borrows the numpy buffers."""
        n = self.n_bodies
        arrs = {
            "body_kind": np.ascontiguousarray(self.body_kind, np.uint8),
            "body_geom": np.ascontiguousarray(self.body_geom, np.float64).reshape(n, 8),
            "coloration_kind": np.ascontiguousarray(self.coloration_kind, np.uint8),
            "color": np.ascontiguousarray(self.color, np.float32).reshape(n, 3),
            "texture_id": np.ascontiguousarray(self.texture_id, np.int32),
            "texture_offset": np.ascontiguousarray(self.texture_offset, np.float32).reshape(n, 2),
            "albedo": np.ascontiguousarray(self.albedo, np.float32),
            "surface_kind": np.ascontiguousarray(self.surface_kind, np.uint8),
            "surface_param": np.ascontiguousarray(self.surface_param, np.float32).reshape(n, 2),
            "light_kind": np.ascontiguousarray(self.light_kind, np.uint8),
            "light_vec": np.ascontiguousarray(self.light_vec, np.float64).reshape(self.n_lights, 3),
            "light_color": np.ascontiguousarray(self.light_color, np.float32).reshape(self.n_lights, 3),
            "light_intensity": np.ascontiguousarray(self.light_intensity, np.float32),
        }
        for name in ("coloration_kind", "texture_id", "albedo", "surface_kind"):
            if arrs[name].shape[0] != n:
                raise SceneError(f"{name} has {arrs[name].shape[0]} rows, expected {n}")
        desc = SceneDesc()
        desc.abi_version = ABI_VERSION
        desc.max_recursion_depth = int(self.max_recursion_depth)
        desc.fov = float(self.fov)
        dc = np.asarray(self.default_color, np.float32)
        desc.default_color[0], desc.default_color[1], desc.default_color[2] = (
            float(dc[0]), float(dc[1]), float(dc[2]))
        desc.n_bodies = n
        desc.n_lights = self.n_lights
        for name, a in arrs.items():
            setattr(desc, name, a.ctypes.data if a.size else None)
        texs = [np.ascontiguousarray(t, np.uint8) for t in self.textures]
        tdescs = (TextureDesc * max(len(texs), 1))()
        for i, t in enumerate(texs):
            if t.ndim != 3 or t.shape[2] not in (3, 4):
                raise SceneError("textures must be (H, W, 3|4) uint8")
            tdescs[i].height, tdescs[i].width, tdescs[i].channels = t.shape
            tdescs[i].pixels = t.ctypes.data
        desc.n_textures = len(texs)
        desc.textures = ctypes.cast(tdescs, ctypes.c_void_p).value if texs else None
        return desc, (arrs, texs, tdescs)


# --------------------------------------------------------------------------- parsing
def parse_color(s: Any) -> np.ndarray:
    """color.rs:114-130: ``#rrggbb`` -> byte / 255.0 in f32."""
    if not isinstance(s, str) or len(s) != 7 or not s.startswith("#"):
        raise SceneError(f"{s} is not a valid color")
    try:
        num = int(s[1:], 16)
    except ValueError:
        raise SceneError(f"{s} is not a valid color") from None
    if num < 0:
        raise SceneError(f"{s} is not a valid color")
    b = np.array([(num >> 16) & 0xFF, (num >> 8) & 0xFF, num & 0xFF], np.float32)
    return b / np.float32(255.0)


def _num(v: Any, what: str) -> float:
    if isinstance(v, bool):
        raise SceneError(f"{what}: expected a number, got a bool")
    if isinstance(v, (int, float)):
        return float(v)
    if isinstance(v, str):  # PyYAML (YAML 1.1) leaves "1e-3" a string; serde_yaml parses it
        try:
            return float(v)
        except ValueError:
            pass
    raise SceneError(f"{what}: expected a number, got {v!r}")


def _f32(v: Any, what: str) -> np.float32:
    # serde_yaml hands f64 to the f32 visitor, which narrows with `as f32`
    return np.float32(_num(v, what))


def _vec3(v: Any, what: str) -> List[float]:
    if isinstance(v, dict):
        try:
            return [_num(v["x"], what), _num(v["y"], what), _num(v["z"], what)]
        except KeyError as e:
            raise SceneError(f"{what}: missing field {e}") from None
    if isinstance(v, (list, tuple)) and len(v) == 3:
        return [_num(c, what) for c in v]
    raise SceneError(f"{what}: expected [x, y, z] or {{x, y, z}}")


def _variant(v: Any, what: str, unit_variants: Sequence[str] = ()):
    """Externally tagged serde enum: ``Name`` (unit) or ``{Name: payload}``."""
    if isinstance(v, str):
        if v in unit_variants:
            return v, None
        raise SceneError(f"{what}: unknown or non-unit variant {v!r}")
    if isinstance(v, dict) and len(v) == 1:
        (k, payload), = v.items()
        return k, payload
    raise SceneError(f"{what}: expected a single-key map naming the variant")


def _need(d: Any, key: str, what: str):
    if not isinstance(d, dict) or key not in d:
        raise SceneError(f"{what}: missing field `{key}`")
    return d[key]


def default_texture_loader(path: str) -> np.ndarray:
    """material.rs:34-47 (`image::open`), decoded by the native host library
    (host/rgh_jpeg.cpp restates jpeg-decoder 0.1.11; host/rgh_png.cpp); RGB8 or RGBA8."""
    from . import host

    return host.open_image(path)


def scene_from_dict(doc: Any, texture_loader: Callable[[str], np.ndarray] = default_texture_loader
                    ) -> SceneData:
    if doc is None:
        doc = {}
    if not isinstance(doc, dict):
        raise SceneError("scene root must be a map")
    known = {"fov", "defaultColor", "maxRecursionDepth", "bodies", "lights"}
    unknown = set(doc) - known
    if unknown:  # scene.rs:12 deny_unknown_fields
        raise SceneError(f"unknown field `{sorted(unknown)[0]}`, expected one of {sorted(known)}")
    sd = SceneData()
    if "fov" in doc:
        sd.fov = _num(doc["fov"], "fov")
    if "defaultColor" in doc:
        sd.default_color = parse_color(doc["defaultColor"])
    if "maxRecursionDepth" in doc:
        v = doc["maxRecursionDepth"]
        if isinstance(v, bool) or not isinstance(v, int) or v < 0 or v > 0xFFFFFFFF:
            raise SceneError("maxRecursionDepth: expected u32")
        sd.max_recursion_depth = v

    bodies = doc.get("bodies") or []
    lights = doc.get("lights") or []
    n = len(bodies)
    sd.body_kind = np.zeros(n, np.uint8)
    sd.body_geom = np.zeros((n, 8), np.float64)
    sd.coloration_kind = np.zeros(n, np.uint8)
    sd.color = np.zeros((n, 3), np.float32)
    sd.texture_id = np.full(n, -1, np.int32)
    sd.texture_offset = np.zeros((n, 2), np.float32)
    sd.albedo = np.zeros(n, np.float32)
    sd.surface_kind = np.zeros(n, np.uint8)
    sd.surface_param = np.zeros((n, 2), np.float32)
    tex_index: Dict[str, int] = {}

    for i, b in enumerate(bodies):
        what = f"bodies[{i}]"
        kind, p = _variant(b, what)
        if kind == "Sphere":
            sd.body_kind[i] = BODY_SPHERE
            sd.body_geom[i, 0:3] = _vec3(_need(p, "center", what), what + ".center")
            sd.body_geom[i, 3] = _num(_need(p, "radius", what), what + ".radius")
        elif kind == "Plane":
            sd.body_kind[i] = BODY_PLANE
            sd.body_geom[i, 0:3] = _vec3(_need(p, "origin", what), what + ".origin")
            sd.body_geom[i, 3:6] = _vec3(_need(p, "normal", what), what + ".normal")
        elif kind == "Disk":
            sd.body_kind[i] = BODY_DISK
            sd.body_geom[i, 0:3] = _vec3(_need(p, "origin", what), what + ".origin")
            sd.body_geom[i, 3:6] = _vec3(_need(p, "normal", what), what + ".normal")
            sd.body_geom[i, 6] = _num(_need(p, "radius", what), what + ".radius")
        elif kind == "AABB":
            sd.body_kind[i] = BODY_AABB
            bounds = _need(p, "bounds", what)
            if not isinstance(bounds, (list, tuple)) or len(bounds) != 2:
                raise SceneError(what + ".bounds: expected two points")
            sd.body_geom[i, 0:3] = _vec3(bounds[0], what + ".bounds[0]")
            sd.body_geom[i, 3:6] = _vec3(bounds[1], what + ".bounds[1]")
        else:
            raise SceneError(f"{what}: unknown variant `{kind}`, expected Sphere, Plane, Disk or AABB")
        m = _need(p, "material", what)
        ckind, cp = _variant(_need(m, "coloration", what), what + ".coloration")
        if ckind == "Color":
            sd.coloration_kind[i] = COLORATION_COLOR
            sd.color[i] = parse_color(cp)
        elif ckind == "Texture":
            sd.coloration_kind[i] = COLORATION_TEXTURE
            path = _need(cp, "image", what + ".Texture")
            if not isinstance(path, str):
                raise SceneError(what + ".Texture.image: expected a string")
            if path not in tex_index:
                try:
                    img = texture_loader(path)
                except Exception as e:  # material.rs:43-46
                    raise SceneError(f"Could not load texture file {path}: {e}") from None
                tex_index[path] = len(sd.textures)
                sd.textures.append(np.ascontiguousarray(img, np.uint8))
                sd.texture_names.append(path)
            sd.texture_id[i] = tex_index[path]
            sd.texture_offset[i, 0] = _f32(_need(cp, "x_offset", what), what + ".x_offset")
            sd.texture_offset[i, 1] = _f32(_need(cp, "y_offset", what), what + ".y_offset")
        else:
            raise SceneError(f"{what}: unknown coloration `{ckind}`")
        sd.albedo[i] = _f32(_need(m, "albedo", what), what + ".albedo")
        skind, sp = _variant(_need(m, "surface", what), what + ".surface", ("Diffuse",))
        if skind == "Diffuse":
            sd.surface_kind[i] = SURFACE_DIFFUSE
        elif skind == "Reflecting":
            sd.surface_kind[i] = SURFACE_REFLECTING
            sd.surface_param[i, 0] = _f32(_need(sp, "reflectivity", what), what + ".reflectivity")
        elif skind == "Refractive":
            sd.surface_kind[i] = SURFACE_REFRACTIVE
            sd.surface_param[i, 0] = _f32(_need(sp, "index", what), what + ".index")
            sd.surface_param[i, 1] = _f32(_need(sp, "transparency", what), what + ".transparency")
        else:
            raise SceneError(f"{what}: unknown surface `{skind}`")

    nl = len(lights)
    sd.light_kind = np.zeros(nl, np.uint8)
    sd.light_vec = np.zeros((nl, 3), np.float64)
    sd.light_color = np.zeros((nl, 3), np.float32)
    sd.light_intensity = np.zeros(nl, np.float32)
    for i, l in enumerate(lights):
        what = f"lights[{i}]"
        kind, p = _variant(l, what)
        if kind == "Directional":
            sd.light_kind[i] = LIGHT_DIRECTIONAL
            sd.light_vec[i] = _vec3(_need(p, "direction", what), what + ".direction")
        elif kind == "Spherical":
            sd.light_kind[i] = LIGHT_SPHERICAL
            sd.light_vec[i] = _vec3(_need(p, "position", what), what + ".position")
        else:
            raise SceneError(f"{what}: unknown variant `{kind}`")
        sd.light_color[i] = parse_color(_need(p, "color", what))
        sd.light_intensity[i] = _f32(_need(p, "intensity", what), what + ".intensity")
    return sd


def parse_scene(yaml_text: str, texture_loader: Callable[[str], np.ndarray] = default_texture_loader
                ) -> SceneData:
    import yaml

    try:
        doc = yaml.safe_load(yaml_text)
    except yaml.YAMLError as e:
        raise SceneError(f"Could not load YAML: {e}") from None
    return scene_from_dict(doc, texture_loader)


def load_scene(path: str, texture_root: Optional[str] = None) -> SceneData:
    """src/main.rs:117-118. Texture paths resolve against ``texture_root`` (default:
    the CWD, as in the reference — material.rs:41-42)."""
    with open(path, "r", encoding="utf-8") as f:
        text = f.read()

    def loader(p: str) -> np.ndarray:
        full = p if (texture_root is None or os.path.isabs(p)) else os.path.join(texture_root, p)
        return default_texture_loader(full)

    return parse_scene(text, loader)
