"""Known-answer tests of individual reference functions, through the oracle's probes.  The
reference has no unit tests for these, so the expectations are derived
from its source."""
import math

import numpy as np
import pytest

from raingun_b200.scene import BODY_AABB, BODY_DISK, BODY_PLANE, BODY_SPHERE, scene_from_dict


def test_texture_wrap_truncates_toward_zero(oracle):
    # material.rs:70-79: ((val * max) as i32) % max, +max if negative
    assert oracle.texture_wrap(0.0, 512) == 0
    assert oracle.texture_wrap(0.999, 512) == 511
    assert oracle.texture_wrap(1.0, 512) == 0
    assert oracle.texture_wrap(1.5, 512) == 256
    assert oracle.texture_wrap(-0.001, 512) == 0          # -0.512 truncates to 0: texel 0 is double width
    assert oracle.texture_wrap(-0.01, 512) == 512 - 5     # -5.12 -> -5 -> 507
    assert oracle.texture_wrap(-1.0, 512) == 0
    assert oracle.texture_wrap(float("nan"), 512) == 0    # `as i32` of NaN is 0
    assert oracle.texture_wrap(1e30, 7) == (2**31 - 1) % 7   # saturating cast


def test_quantise_truncates_and_saturates(oracle):
    # color.rs:32-37 `(c * 255.0) as u8`; Display floors the same way (color.rs:187-191)
    assert oracle.quantise(1.0) == 255 and oracle.quantise(0.5) == 127 and oracle.quantise(0.0) == 0
    assert oracle.quantise(0.9999) == 254 and oracle.quantise(2.0) == 255 and oracle.quantise(-1.0) == 0
    assert oracle.quantise(float("nan")) == 0


def test_fresnel_is_the_normal_incidence_value(oracle):
    # rendering.rs:174-200 uses cos_i = |cos_t|, so kr = ((eta_t - eta_i)/(eta_t + eta_i))^2 unless TIR
    n = [0.0, 0.0, 1.0]
    idx = np.float32(1.33)
    expect = ((float(idx) - 1.0) / (float(idx) + 1.0)) ** 2
    for inc in ([0.0, 0.0, -1.0], [0.6, 0.0, -0.8], [0.0, 0.0, 1.0]):
        assert oracle.fresnel(inc, n, idx) == pytest.approx(expect, rel=1e-12)
    # inside the body at a grazing angle: total internal reflection -> exactly 1
    assert oracle.fresnel([math.sqrt(1 - 0.1**2), 0.0, 0.1], n, idx) == 1.0


def test_sphere_intersection_cases(oracle):
    g = [0.0, 0.0, -5.0, 1.0]
    assert oracle.intersect(BODY_SPHERE, g, [0, 0, 0], [0, 0, -1]) == 4.0          # outside: near root
    assert oracle.intersect(BODY_SPHERE, g, [0, 0, -5], [0, 0, -1]) == 1.0         # inside: far root
    assert oracle.intersect(BODY_SPHERE, g, [0, 0, -7], [0, 0, -1]) is None        # behind
    assert oracle.intersect(BODY_SPHERE, g, [0, 1.0000001, 0], [0, 0, -1]) is None  # misses
    # |d| != 1 is never renormalised (bodies.rs:93-95 assumes a unit direction)
    assert oracle.intersect(BODY_SPHERE, g, [0, 0, 0], [0, 0, -2]) == pytest.approx(10.0 - math.sqrt(1 - 25 + 100))


def test_plane_is_one_sided_and_disk_uses_strict_radius(oracle):
    plane = [0.0, -2.0, 0.0, 0.0, -1.0, 0.0]
    assert oracle.intersect(BODY_PLANE, plane, [0, 0, 0], [0, -1, 0]) == 2.0   # along +normal: hit
    assert oracle.intersect(BODY_PLANE, plane, [0, -4, 0], [0, 1, 0]) is None  # against the normal: never
    assert oracle.intersect(BODY_PLANE, plane, [0, 0, 0], [1, -1e-7, 0]) is None  # denominator <= 1e-6
    disk = [0.0, 0.0, -3.0, 0.0, 0.0, -1.0, 2.0]
    assert oracle.intersect(BODY_DISK, disk, [1.5, 0, 0], [0, 0, -1]) == 3.0
    assert oracle.intersect(BODY_DISK, disk, [2.0, 0, 0], [0, 0, -1]) is None   # sqrt(d2) < radius is strict


def test_aabb_returns_tmin_or_tmax(oracle):
    box = [-1.0, -1.0, -6.0, 1.0, 1.0, -4.0]
    assert oracle.intersect(BODY_AABB, box, [0, 0, 0], [0, 0, -1]) == 4.0
    assert oracle.intersect(BODY_AABB, box, [0, 0, -5], [0, 0, -1]) == 1.0     # origin inside: tmax
    assert oracle.intersect(BODY_AABB, box, [0, 0, -7], [0, 0, -1]) is None
    assert oracle.intersect(BODY_AABB, box, [3, 0, 0], [0, 0, -1]) is None


def _two_spheres(depth):
    mat = lambda surface: {"coloration": {"Color": "#ffffff"}, "albedo": 0.5, "surface": surface}
    return scene_from_dict({
        "maxRecursionDepth": depth, "defaultColor": "#336699",
        "lights": [{"Directional": {"direction": [0, -1, -1], "color": "#ffffff", "intensity": 3.0}}],
        "bodies": [{"Sphere": {"center": [0, 0, -4], "radius": 1.0, "material": mat({"Reflecting": {"reflectivity": 0.5}})}},
                   {"Sphere": {"center": [0, 0, -4], "radius": 1.0, "material": mat("Diffuse")}}]})


def test_first_minimum_tie_break_and_depth_semantics(oracle):
    """Two coincident spheres: min_by keeps the FIRST of equal minima (scene.rs:34-39), so the
    reflecting one (index 0) is what is seen.  Depth 0 still traces the primary ray but returns
    default_color for the reflection without tracing it (rendering.rs:122-124)."""
    _, st0, _ = oracle.render(_two_spheres(0), 64, 48)
    assert st0.rays_reflection == 0 and st0.rays_primary == 64 * 48 and st0.rays_shadow > 0
    _, st1, _ = oracle.render(_two_spheres(1), 64, 48)
    assert st1.rays_reflection == 0                       # depth+1 = 1 >= max 1: not traced
    _, st2, _ = oracle.render(_two_spheres(2), 64, 48)
    assert st2.rays_reflection > 0                        # only possible if body 0 won the tie
    assert st2.rays_shadow == st0.rays_shadow + 0 or st2.rays_shadow >= st0.rays_shadow


def test_portrait_is_rejected(oracle):
    with pytest.raises(RuntimeError):
        oracle.render(_two_spheres(1), 48, 64)            # ray.rs:42 assert!(width >= height)
