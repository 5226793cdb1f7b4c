"""Host-side scene ingestion: the reference's serde schema (scene.rs:11-31, bodies.rs:13-47,
lights.rs:8-26, material.rs:7-54, color.rs:114-160) and src/main.rs's depth override."""
import numpy as np
import pytest

from raingun_b200.examples import bundled_texture_loader, example_scene, example_yaml
from raingun_b200.scene import SceneError, parse_color, parse_scene
from raingun_b200.synth import SPECS, SplitMix64, make_scene, make_scene_doc, to_yaml


def test_color_parsing_matches_reference_unit_tests():
    # color.rs:162-185 it_parses_strings
    assert parse_color("#000000").tolist() == [0.0, 0.0, 0.0]
    assert parse_color("#ffffff").tolist() == [1.0, 1.0, 1.0]
    c = parse_color("#ff7f11")
    assert c.dtype == np.float32
    assert c[0] == np.float32(1.0) and c[1] == np.float32(0.498039216) and c[2] == np.float32(0.066666667)
    for bad in ("#fff", "ffffff0", "#gggggg", 7, None, "#12345"):
        with pytest.raises(SceneError):
            parse_color(bad)


def test_example_scenes_flatten():
    t1 = example_scene("test1")
    assert t1.n_bodies == 6 and t1.n_lights == 3 and t1.max_recursion_depth == 10 and t1.fov == 90.0
    assert t1.body_kind.tolist() == [1, 1, 0, 0, 0, 0]
    assert t1.surface_kind.tolist() == [0, 0, 0, 1, 0, 2]       # `Diffuse` and `Diffuse:` both accepted
    assert len(t1.textures) == 2 and t1.textures[0].shape == (1500, 1500, 3) and t1.textures[1].shape == (1024, 2048, 3)
    assert t1.texture_offset[2, 0] == np.float32(0.9)
    assert t1.surface_param[5, 0] == np.float32(1.33) and float(t1.surface_param[5, 0]) == 1.3300000429153442
    assert t1.light_vec[0].tolist() == [0.4, -1.0, -0.9]        # map form {x,y,z}
    assert t1.light_vec[1].tolist() == [-6.0, 3.2, -5.0]        # sequence form
    t2 = example_scene("test2")
    assert t2.body_kind.tolist() == [1, 2, 2, 3]
    assert t2.body_geom[2, 3:7].tolist() == [0.0, 0.2, -0.7, 4.0]   # un-normalised disk normal is kept as given
    assert t2.body_geom[3, :6].tolist() == [2.0, -2.0, -3.2, 2.8, -0.4, -3.0]


def test_defaults_and_strictness():
    s = parse_scene("---\n{}\n")
    assert s.fov == 90.0 and s.max_recursion_depth == 10 and s.default_color.tolist() == [0, 0, 0] and s.n_bodies == 0
    with pytest.raises(SceneError):                      # scene.rs:12 deny_unknown_fields
        parse_scene("fov: 60\nbogus: 1\n")
    with pytest.raises(SceneError):                      # snake_case root keys are unknown fields
        parse_scene("default_color: '#ffffff'\n")
    with pytest.raises(SceneError):
        parse_scene("bodies:\n  - Cone: {}\n")
    with pytest.raises(SceneError):                      # missing required field
        parse_scene("bodies:\n  - Sphere: {center: [0,0,0], material: {coloration: {Color: '#ffffff'}, albedo: 1, surface: Diffuse}}\n")
    with pytest.raises(SceneError):                      # texture that cannot be loaded
        parse_scene(example_yaml("test3"), lambda p: (_ for _ in ()).throw(IOError("nope")))
    s = parse_scene("maxRecursionDepth: 3\nfov: 45.5\ndefaultColor: '#0000ff'\n")
    assert s.max_recursion_depth == 3 and s.fov == 45.5 and s.default_color.tolist() == [0, 0, 1]


def test_depth_limit_only_lowers():
    s = example_scene("test2")
    assert s.with_max_depth_limit(4).max_recursion_depth == 4      # --draft, main.rs:74-75
    assert s.with_max_depth_limit(50).max_recursion_depth == 10
    assert s.with_max_depth_limit(None).max_recursion_depth == 10


def test_splitmix_known_answers():
    r = SplitMix64(0)
    assert [r.next_u64() for _ in range(3)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]


def test_synthetic_scenes_are_deterministic_and_round_trip_yaml():
    doc = make_scene_doc(SPECS["C4"], spheres=50)
    assert doc == make_scene_doc(SPECS["C4"], spheres=50)
    a, spec = make_scene("C4", spheres=50)
    b = parse_scene(to_yaml(doc))                         # through YAML text, as the reference CLI would read it
    assert spec.width == 3840 and a.n_bodies == 51 and a.body_kind[0] == 1
    for f in ("body_kind", "body_geom", "color", "albedo", "surface_kind", "surface_param", "light_vec", "light_intensity"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    kinds = set(a.surface_kind[1:].tolist())
    assert kinds == {0, 1, 2}
    c5, _ = make_scene("C5", spheres=9, texture_loader=bundled_texture_loader)
    assert c5.coloration_kind.tolist() == [1, 1, 0, 0, 0, 1, 0, 0, 0, 1] and len(c5.textures) == 2
    full = make_scene_doc(SPECS["C3"])
    assert len(full["bodies"]) == 1001 and full["maxRecursionDepth"] == 4 and len(full["lights"]) == 3


def test_scene_constructors_use_the_native_loader(monkeypatch, tmp_path):
    """Scene.from_yaml / from_yaml_file parse with libraingun_host (no PyYAML on the product path);
    without a GPU the upload then fails loudly with RG_E_CUDA - there is no CPU fallback."""
    import raingun_b200 as rg
    from raingun_b200 import host, scene as pyscene
    from raingun_b200.examples import bundled_texture_loader, example_yaml

    calls = []
    real = host.parse_scene
    monkeypatch.setattr(host, "parse_scene", lambda *a, **k: calls.append(1) or real(*a, **k))
    monkeypatch.setattr(pyscene, "parse_scene", lambda *a, **k: (_ for _ in ()).throw(AssertionError("python loader used")))
    p = tmp_path / "t2.yml"
    p.write_text(example_yaml("test2"))
    if rg.device_count() == 0:
        for make in (lambda: rg.Scene.from_yaml(example_yaml("test1"), texture_loader=bundled_texture_loader, max_depth=4),
                     lambda: rg.Scene.from_yaml_file(str(p))):
            with pytest.raises(rg.RaingunError) as e:
                make()
            assert e.value.code == rg._native.E_CUDA
    else:
        with rg.Scene.from_yaml_file(str(p)) as sc:
            assert sc.render_image(64, 48).shape == (48, 64, 4)
    assert len(calls) >= 1
