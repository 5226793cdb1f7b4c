"""The native host library (include/raingun_host.h, raingun_b200/host/): C++ YAML scene loader,
JPEG / PNG codecs and CLI option mapping — SURVEY.md section 8(f) row 3.  CPU-only."""
import ctypes
import io
import os
import subprocess

import numpy as np
import pytest

from raingun_b200 import host
from raingun_b200 import scene as pyscene
from raingun_b200.examples import _bundle, bundled_texture_loader, example_yaml
from raingun_b200.synth import make_scene, make_scene_doc, to_yaml

FIELDS = ("body_kind", "body_geom", "coloration_kind", "color", "texture_id", "texture_offset", "albedo",
          "surface_kind", "surface_param", "light_kind", "light_vec", "light_color", "light_intensity")


def assert_same_scene(a, b):
    assert a.fov == b.fov and a.max_recursion_depth == b.max_recursion_depth
    assert np.array_equal(np.asarray(a.default_color, np.float32), np.asarray(b.default_color, np.float32))
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and x.shape == y.shape, f
        assert x.tobytes() == y.tobytes(), f  # bit-equal, NaN-safe
    assert a.texture_names == b.texture_names
    assert len(a.textures) == len(b.textures)
    for x, y in zip(a.textures, b.textures):
        assert np.array_equal(x, y)


def test_library_exports_every_declared_symbol():
    L = host.lib()
    hdr = open(os.path.join(os.path.dirname(host.LIB_PATH), "..", "include", "raingun_host.h")).read()
    import re
    declared = set(re.findall(r"\b(rgh_[a-z0-9_]+)\s*\(", hdr)) - {"rgh_texture_cb"}
    assert declared == set(host.EXPORTS)
    for name in host.EXPORTS:
        assert hasattr(L, name), name


# ------------------------------------------------------------------------------------ YAML / schema
@pytest.mark.parametrize("name", ["test1", "test2", "test3"])
def test_native_loader_equals_python_loader_on_examples(name):
    text = example_yaml(name)
    assert_same_scene(host.parse_scene(text, bundled_texture_loader), pyscene.parse_scene(text, bundled_texture_loader))


@pytest.mark.parametrize("cfg,spheres", [("C3", 200), ("C4", 500), ("C5", 300)])
def test_native_loader_equals_python_loader_on_synthetic_yaml(cfg, spheres):
    data, spec = make_scene(cfg, spheres=spheres, texture_loader=bundled_texture_loader)
    text = to_yaml(make_scene_doc(spec, spheres))  # what the reference CLI would be given
    native = host.parse_scene(text, bundled_texture_loader)
    assert_same_scene(native, data)
    assert native.n_bodies == spheres + 1


@pytest.mark.parametrize("flow", [False, True, None], ids=["block", "flow", "mixed"])
@pytest.mark.parametrize("width", [80, 30, 100000])
def test_yaml_styles_emitted_by_another_yaml_library(flow, width):
    """The same scene document written by PyYAML in block style, in flow style (one long line or
    wrapped at 30 columns, i.e. flow collections continued over many lines) and mixed."""
    import yaml

    data, spec = make_scene("C5", spheres=60, texture_loader=bundled_texture_loader)
    doc = make_scene_doc(spec, 60)
    text = yaml.safe_dump(doc, sort_keys=False, default_flow_style=flow, width=width)
    assert_same_scene(host.parse_scene(text, bundled_texture_loader), data)
    text2 = "%YAML 1.1\n--- # scene\n" + yaml.safe_dump(doc, sort_keys=True, default_flow_style=flow, width=width, indent=6) + "...\n"
    assert_same_scene(host.parse_scene(text2, bundled_texture_loader), data)


def test_yaml_forms_the_reference_schema_allows():
    text = """
# comment line
fov: 60   # integer where f64 is expected
maxRecursionDepth: 3
defaultColor: '#102030'
lights:
- Spherical: {position: {x: 1, y: 2.5, z: -3e0}, color: "#ffffff", intensity: 1e3}
- Directional:
    direction: [0.0, -1.0,
                0.0]
    color: "#ff0000"
    intensity: +2
bodies:
  - Disk:
      origin: [0, 0, -5]
      normal: {x: 0, y: 0, z: -1}
      radius: 2
      ignored_extra_key: 1      # only the root denies unknown fields (scene.rs:12)
      material: {coloration: {Color: "#00ff00"}, albedo: 0.5, surface: Diffuse}
  - AABB:
      bounds: [[-1, -1, -6], {x: 1, y: 1, z: -4}]
      material:
        coloration: {Color: "#0000ff"}
        albedo: .25
        surface: {Refractive: {index: 1.33, transparency: 0.8}}
  - Sphere:
      center: [0, 0, -3]
      radius: 1.
      material:
        coloration:
          Color: "#+00aBc"
        albedo: 1
        surface:
          Reflecting:
            reflectivity: 0.75
"""
    sd = host.parse_scene(text)
    assert sd.fov == 60.0 and sd.max_recursion_depth == 3
    assert np.allclose(sd.default_color, np.array([0x10, 0x20, 0x30], np.float32) / np.float32(255))
    assert list(sd.light_kind) == [pyscene.LIGHT_SPHERICAL, pyscene.LIGHT_DIRECTIONAL]
    assert sd.light_vec.tolist() == [[1.0, 2.5, -3.0], [0.0, -1.0, 0.0]]
    assert sd.light_intensity.tolist() == [1000.0, 2.0]
    assert list(sd.body_kind) == [pyscene.BODY_DISK, pyscene.BODY_AABB, pyscene.BODY_SPHERE]
    assert sd.body_geom[0].tolist() == [0, 0, -5, 0, 0, -1, 2, 0]
    assert sd.body_geom[1].tolist() == [-1, -1, -6, 1, 1, -4, 0, 0]
    assert sd.surface_param[1].tolist() == [float(np.float32(1.33)), float(np.float32(0.8))]  # f64 -> `as f32`
    assert sd.albedo.tolist() == [0.5, 0.25, 1.0]
    assert np.array_equal(sd.color[2], np.array([0x00, 0x0a, 0xbc], np.float32) / np.float32(255))
    assert_same_scene(sd, pyscene.parse_scene(text.replace('"#+00aBc"', '"#000abc"').replace("intensity: +2", "intensity: 2")))


def test_yaml_anchors_and_aliases():
    """Hand-written scenes share materials through anchors; serde_yaml resolves them, so must we —
    in block position, inline, in flow collections, and for whole sequence items."""
    text = """
defaultColor: &bg "#102030"
lights:
  - &sun
    Directional:
      direction: &down [0.0, -1.0, 0.0]
      color: *bg
      intensity: &one 1
  - *sun
  - Spherical: {position: *down, color: "#ffffff", intensity: *one}
bodies:
  - Sphere:
      center: [0, 0, -3]
      radius: *one
      material: &glass
        coloration: {Color: &white "#ffffff"}
        albedo: 0.5
        surface: &refr {Refractive: {index: 1.5, transparency: 0.9}}
  - Sphere: {center: [2, 0, -4], radius: 0.5, material: *glass}
  - Plane:
      origin: [0, -2, 0]
      normal: *down
      material:
        coloration:
          Color: *white
        albedo: *one
        surface: *refr
  - &ball
    Sphere: {center: [-2, 0, -4], radius: 0.25, material: {coloration: {Color: *bg}, albedo: 1, surface: &d Diffuse}}
  - *ball
  - &kind Sphere: {center: [-3, 0, -4], radius: 0.25, material: {coloration: {Color: *bg}, albedo: 1, surface: *d}}
"""
    import yaml

    native = host.parse_scene(text)
    assert_same_scene(native, pyscene.scene_from_dict(yaml.safe_load(text)))
    assert native.n_bodies == 6 and native.n_lights == 3
    assert native.surface_kind.tolist() == [2, 2, 2, 0, 0, 0] and native.body_kind.tolist() == [0, 0, 1, 0, 0, 0]
    assert np.array_equal(native.body_geom[3], native.body_geom[4])


def test_empty_document_is_the_default_scene():
    for text in ("", "---\n", "# nothing\n"):
        sd = host.parse_scene(text)
        assert sd.fov == 90.0 and sd.max_recursion_depth == 10 and sd.n_bodies == 0 and sd.n_lights == 0
        assert sd.default_color.tolist() == [0.0, 0.0, 0.0]


@pytest.mark.parametrize("text,code,needle", [
    ("fov: 90\nfoo: 1\n", host.E_SCHEMA, "unknown field `foo`"),                       # scene.rs:12
    ("defaultColor: \"#12345\"\n", host.E_SCHEMA, "is not a valid color"),             # color.rs:129
    ("defaultColor: \"#12345g\"\n", host.E_SCHEMA, "is not a valid color"),
    ("defaultColor: 7\n", host.E_SCHEMA, "expected a string of a simple hex color"),   # color.rs:138
    ("fov: \"90\"\n", host.E_SCHEMA, "expected f64"),                                  # quoted = string
    ("maxRecursionDepth: -1\n", host.E_SCHEMA, "expected u32"),
    ("bodies:\n  - Cube: {}\n", host.E_SCHEMA, "unknown variant `Cube`"),
    ("bodies:\n  - Sphere:\n      center: [0, 0, -3]\n", host.E_SCHEMA, "missing field `radius`"),
    ("bodies:\n  - Sphere:\n      center: [0, 0]\n      radius: 1\n", host.E_SCHEMA, "expected [x, y, z]"),
    ("lights:\n  - Spherical: {position: [0,0,0], color: \"#ffffff\"}\n", host.E_SCHEMA, "missing field `intensity`"),
    ("bodies:\n  - Sphere: {center: [0,0,0], radius: 1, material: {coloration: {Texture: {image: \"nope.jpg\", "
     "x_offset: 0, y_offset: 0}}, albedo: 1, surface: Diffuse}}\n", host.E_SCHEMA, "Could not load texture file nope.jpg"),
    ("bodies:\n  - Sphere: {center: [0,0,0], radius: 1, material: {coloration: {Color: \"#ffffff\"}, albedo: 1, "
     "surface: Reflecting}}\n", host.E_SCHEMA, "Reflecting"),                           # not a unit variant
    ("a: [1, 2\n", host.E_FORMAT, "unterminated flow"),
    ("a: *nope\n", host.E_FORMAT, "unknown anchor"),
    ("a: !!str 1\n", host.E_UNSUPPORTED, "tags"),
    ("a: |\n  text\n", host.E_UNSUPPORTED, "block scalars"),
    ("a: 1\n---\nb: 2\n", host.E_UNSUPPORTED, "multi-document"),
    ("\tfov: 1\n", host.E_FORMAT, "tab"),
])
def test_errors_mirror_serde(text, code, needle):
    with pytest.raises(host.HostError) as e:
        host.parse_scene(text)
    assert e.value.code == code, str(e.value)
    assert needle in str(e.value), str(e.value)
    assert str(e.value).startswith("Could not load YAML")  # src/main.rs:118


def test_depth_limit_only_lowers():
    text = "maxRecursionDepth: 6\n"
    assert host.parse_scene(text, max_depth_limit=4).max_recursion_depth == 4  # --draft, main.rs:119-123
    assert host.parse_scene(text, max_depth_limit=9).max_recursion_depth == 6


# ------------------------------------------------------------------------------------ codecs
@pytest.mark.parametrize("key,shape", [
    ("texture/./textures/clay-ground-seamless.jpg", (1500, 1500, 3)),        # progressive, 4:4:4, Adobe
    ("texture/./textures/land_ocean_ice_cloud_2048.jpg", (1024, 2048, 3)),   # baseline, 4:4:4, Adobe
    ("texture/textures/tile1/color.jpg", (512, 512, 3)),                     # baseline, 4:2:0 (H2V2)
])
def test_jpeg_decoder_agrees_with_libjpeg_within_idct_tolerance(key, shape):
    """Entropy decoding is exact; IDCT / upsampling / colour conversion legitimately differ between
    decoders by a few LSB.  (Equality with the REFERENCE's decoder is pinned by test_oracle_golden.)"""
    from PIL import Image

    mine = host.decode_jpeg(_bundle()[key])
    assert mine.shape == shape
    theirs = np.asarray(Image.open(io.BytesIO(_bundle()[key])).convert("RGB"))
    d = np.abs(mine.astype(np.int32) - theirs.astype(np.int32))
    assert d.max() <= 4 and d.mean() < 0.1


def test_jpeg_variants_roundtrip_through_pillow_encoder():
    """Grey, 4:2:2 (H2V1), restart intervals, progressive+subsampled, odd sizes."""
    from PIL import Image

    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    base = np.asarray(Image.fromarray(base).resize((101, 67), Image.BILINEAR))  # smooth-ish, odd size
    cases = [dict(subsampling=0), dict(subsampling=1), dict(subsampling=2), dict(subsampling=2, progressive=True),
             dict(subsampling=0, progressive=True), dict(subsampling=2, restart_marker_blocks=3),
             dict(subsampling=1, progressive=True, restart_marker_rows=1), dict(grey=True), dict(grey=True, progressive=True)]
    for kw in cases:
        grey = kw.pop("grey", False)
        im = Image.fromarray(base).convert("L") if grey else Image.fromarray(base)
        buf = io.BytesIO()
        im.save(buf, "JPEG", quality=90, **kw)
        mine = host.decode_jpeg(buf.getvalue())
        theirs = np.asarray(Image.open(io.BytesIO(buf.getvalue())))
        theirs = theirs[..., None] if theirs.ndim == 2 else theirs
        assert mine.shape == theirs.shape, kw
        d = np.abs(mine.astype(np.int32) - theirs.astype(np.int32))
        # fancy-upsampling edge handling differs by design (stb-style vs libjpeg): a loose bound
        assert d.mean() < 1.0 and d.max() <= 24, (kw, d.max(), d.mean())


def test_jpeg_rejects_garbage():
    good = _bundle()["texture/textures/tile1/color.jpg"]
    for bad, code in ((b"", host.E_FORMAT), (b"\x89PNG", host.E_FORMAT), (good[:200], host.E_FORMAT)):
        with pytest.raises(host.HostError) as e:
            host.decode_jpeg(bad)
        assert e.value.code == code


def test_png_roundtrip_and_against_pillow(tmp_path):
    from PIL import Image

    rng = np.random.default_rng(9)
    for ch in (1, 3, 4):
        a = rng.integers(0, 256, (33, 47, ch), dtype=np.uint8)
        a[5:20, 3:30] = 77  # something compressible
        data = host.encode_png(a)
        back = host.decode_png(data)
        want = np.repeat(a, 3, axis=2) if ch == 1 else a
        assert np.array_equal(back, want)
        pil = np.asarray(Image.open(io.BytesIO(data)))
        assert np.array_equal(pil if pil.ndim == 3 else pil[..., None], a)
    # decode what another encoder wrote: palette, 1-bit grey, interlaced RGBA, grey+alpha
    img = Image.fromarray(rng.integers(0, 256, (19, 23, 3), dtype=np.uint8))
    for mode, kw in (("P", {}), ("1", {}), ("RGBA", {}), ("LA", {}), ("L", {}), ("RGB", {"optimize": True})):
        im = img.convert(mode)
        buf = io.BytesIO()
        im.save(buf, "PNG", **kw)
        mine = host.decode_png(buf.getvalue())
        want = np.asarray(im.convert("RGBA" if mine.shape[2] == 4 else "RGB"))
        assert np.array_equal(mine, want), mode
    p = str(tmp_path / "x.png")
    host.save_png(p, a)
    assert np.array_equal(host.open_image(p), a)


def test_golden_pngs_decode_like_pillow():
    from PIL import Image

    for name in ("test1", "test2", "test3"):
        data = _bundle()[f"golden/{name}.png"]
        mine = host.decode_png(data)
        theirs = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA" if mine.shape[2] == 4 else "RGB"))
        assert np.array_equal(mine, theirs)


# ------------------------------------------------------------------------------------ CLI (src/main.rs:143-178)
def test_cli_resolution_arguments_like_the_reference_tests():
    P = host.cli_parse
    assert P(["x", "file"])[:3] == (800, 600, None)
    assert P(["x", "--width", "640", "--height", "480", "file"])[:2] == (640, 480)
    assert P(["x", "--hd", "file"])[:2] == (1920, 1080)
    assert P(["x", "--hd", "--4k", "file"])[:2] == (3840, 2160)
    assert P(["x", "--hd", "--width", "2000", "file"])[:2] == (2000, 1080)
    # it_parses_draft_argument
    assert P(["x", "--hd", "--width", "2000", "--draft", "file"])[:3] == (800, 600, 4)
    # forms clap accepts
    assert P(["x", "-w", "320", "-h240", "--output=o.png", "scene.yml"]) == (320, 240, None, False, "scene.yml", "o.png")
    assert P(["x", "--4k", "--hd", "--preview", "a/b.scene.yml"]) == (1920, 1080, None, True, "a/b.scene.yml", "a/b.scene.png")
    assert P(["x", "noext"])[5] == "noext.png"
    assert P(["x", "dir.d/.hidden"])[5] == "dir.d/.hidden.png"


@pytest.mark.parametrize("argv,needle", [
    (["x"], "required arguments"), (["x", "--width", "abc", "f"], "Could not parse width"),
    (["x", "--height", "-3", "f"], "Could not parse height"), (["x", "--bogus", "f"], "wasn't expected"),
    (["x", "a", "b"], "wasn't expected"), (["x", "--width"], "requires a value"), (["x", ".."], "Could not guess output filename"),
])
def test_cli_usage_errors(argv, needle):
    with pytest.raises(host.HostError) as e:
        host.cli_parse(argv)
    assert e.value.code == host.E_USAGE and needle in str(e.value)


def test_cli_binary_fails_loudly_without_a_gpu(tmp_path):
    """The CLI has no CPU fallback: on a box without an sm_100 device the upload fails, status 101
    (the reference's panic status); usage errors exit 1 like clap."""
    if not os.path.exists(host.CLI_PATH):
        pytest.skip("CLI not built")
    from raingun_b200 import device_count

    scene = tmp_path / "s.yml"
    scene.write_text(example_yaml("test2"))
    r = subprocess.run([host.CLI_PATH], capture_output=True, text=True)
    assert r.returncode == 1 and "USAGE" in r.stderr
    r = subprocess.run([host.CLI_PATH, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "USAGE" in r.stdout and "--draft" in r.stdout
    r = subprocess.run([host.CLI_PATH, "-V"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "raingun 0.1.0"   # the reference's crate version
    if device_count() == 0:
        r = subprocess.run([host.CLI_PATH, "--draft", str(scene)], capture_output=True, text=True)
        assert r.returncode == 101 and "Could not upload the scene" in r.stderr
        assert not (tmp_path / "s.png").exists()


# ------------------------------------------------------------------------------------ hostile input
@pytest.mark.timeout(120)
def test_truncated_and_corrupt_images_fail_cleanly_and_quickly():
    """The decoders read untrusted files.  A JPEG cut off inside a scan once sent the marker loop back
    to an earlier 0xFF for ever; corrupt dimensions must not allocate the machine away; nothing may
    crash or hang (the C++ sources are also fuzzed under ASan/UBSan)."""
    import struct
    import zlib

    rng = np.random.default_rng(3)
    for key in ("texture/textures/tile1/color.jpg", "texture/./textures/land_ocean_ice_cloud_2048.jpg",
                "texture/./textures/clay-ground-seamless.jpg"):
        data = _bundle()[key]
        cuts = sorted(set([3, 20, 200, 700, len(data) // 3, len(data) // 2, len(data) - 2] +
                          [int(x) for x in rng.integers(2, len(data), 6)]))
        for cut in cuts:
            try:
                img = host.decode_jpeg(data[:cut])       # a prefix may still decode (missing rows are grey)
                assert img.ndim == 3
            except host.HostError as e:
                assert e.code in (host.E_FORMAT, host.E_UNSUPPORTED)
        # frame header claiming 65535 x 65535
        i = 2
        while data[i + 1] not in (0xC0, 0xC1, 0xC2):        # walk the marker segments up to the frame header
            i += 2 + ((data[i + 2] << 8) | data[i + 3])
        huge = data[:i + 5] + b"\xff\xff\xff\xff" + data[i + 9:]
        with pytest.raises(host.HostError):
            host.decode_jpeg(huge)
    # PNG: corrupt IDAT, truncated file, absurd dimensions with a valid CRC
    png = host.encode_png(rng.integers(0, 256, (20, 30, 3), dtype=np.uint8))
    for bad in (png[:40], png[:-20], png[:60] + b"\x00" * 10 + png[70:]):
        with pytest.raises(host.HostError):
            host.decode_png(bad)
    ihdr = struct.pack(">IIBBBBB", 0x7fffffff, 0x7fffffff, 8, 6, 0, 0, 0)
    big = png[:8] + struct.pack(">I", 13) + b"IHDR" + ihdr + struct.pack(">I", zlib.crc32(b"IHDR" + ihdr)) + png[33:]
    with pytest.raises(host.HostError):
        host.decode_png(big)


def test_yaml_nesting_is_bounded():
    with pytest.raises(host.HostError) as e:
        host.parse_scene("a: " + "[" * 5000 + "]" * 5000 + "\n")
    assert e.value.code == host.E_UNSUPPORTED and "nested" in str(e.value)
    with pytest.raises(host.HostError):
        host.parse_scene("".join(" " * i + "k:\n" for i in range(400)))


def test_bmp_tga_pnm_decoders_against_pillow(tmp_path):
    """The simple formats `image::open` also accepts (material.rs:42): decode what Pillow writes, in
    every variant it can write, and compare with Pillow's own reading of the same bytes."""
    from PIL import Image

    rng = np.random.default_rng(21)
    base = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    base[10:20, 5:40] = (200, 30, 90)          # runs, so that RLE has something to do
    rgb = Image.fromarray(base)

    def check(fmt, im, decode, **kw):
        buf = io.BytesIO()
        im.save(buf, fmt, **kw)
        data = buf.getvalue()
        mine = decode(data)
        back = Image.open(io.BytesIO(data))
        want = np.asarray(back.convert({1: "L", 3: "RGB", 4: "RGBA"}[mine.shape[2]]))
        want = want[..., None] if want.ndim == 2 else want
        assert mine.shape == want.shape, (fmt, im.mode, kw, mine.shape, want.shape)
        assert np.array_equal(mine, want), (fmt, im.mode, kw, int(np.abs(mine.astype(int) - want.astype(int)).max()))
        return data

    for mode in ("RGB", "RGBA", "P", "L", "1"):
        check("BMP", rgb.convert(mode), host.decode_bmp)
    for mode in ("RGB", "RGBA", "L", "P"):
        for rle in (False, True):
            check("TGA", rgb.convert(mode), host.decode_tga, compression="tga_rle" if rle else None)
    check("PPM", rgb, host.decode_pnm)
    check("PPM", rgb.convert("L"), host.decode_pnm)
    check("PPM", rgb.convert("1"), host.decode_pnm)
    # ASCII variants and a 16-bit maxval, which Pillow does not write
    a = base[:4, :5]
    p3 = ("P3\n# comment\n5 4\n255\n" + " ".join(str(int(v)) for v in a.reshape(-1)) + "\n").encode()
    assert np.array_equal(host.decode_pnm(p3), a)
    p2 = ("P2 5 4 1023\n" + "\n".join(" ".join(str(int(v) * 4) for v in row) for row in a[..., 0]) + "\n").encode()
    assert np.abs(host.decode_pnm(p2)[..., 0].astype(int) - a[..., 0]).max() <= 1
    p1 = b"P1\n3 2\n1 0 1\n0 1 0\n"
    assert host.decode_pnm(p1)[..., 0].tolist() == [[0, 255, 0], [255, 0, 255]]
    # image::open picks the decoder by extension; a texture in one of these formats loads into a scene
    for ext, fmt in (("bmp", "BMP"), ("tga", "TGA"), ("ppm", "PPM")):
        path = tmp_path / f"tex.{ext}"
        rgb.save(path, fmt)
        assert np.array_equal(host.open_image(str(path)), base)
    text = ("bodies:\n  - Sphere: {center: [0, 0, -3], radius: 1, material: {coloration: {Texture: {image: tex.bmp, "
            "x_offset: 0, y_offset: 0}}, albedo: 1, surface: Diffuse}}\n")
    sd = host.parse_scene(text, texture_root=str(tmp_path))
    assert np.array_equal(sd.textures[0], base)
    for ext in ("webp", "tiff", "ico"):
        (tmp_path / f"t.{ext}").write_bytes(b"xx")
        with pytest.raises(host.HostError) as e:
            host.open_image(str(tmp_path / f"t.{ext}"))
        assert e.value.code == host.E_UNSUPPORTED
    for decode in (host.decode_bmp, host.decode_tga, host.decode_pnm):
        for bad in (b"", b"BM", b"P6 4 4 255\nxx", b"\x00" * 17, b"\x00\x00\x02" + b"\x00" * 9 + b"\x10\x00\x10\x00\x18\x00"):
            try:
                decode(bad)
            except host.HostError as e:
                assert e.code in (host.E_FORMAT, host.E_UNSUPPORTED, host.E_INVALID)


def test_scene_load_from_a_directory_laid_out_like_the_reference(tmp_path):
    """rgh_scene_load + rgh_image_open by path (what the CLI does): the example scenes and their JPEG
    textures written to disk load to exactly what the in-memory route produces."""
    for key, blob in _bundle().items():
        kind, rel = key.split("/", 1)
        if kind == "golden":
            continue
        path = tmp_path / ("examples/" + rel if kind == "scene" else rel)
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_bytes(blob)
    for key in ("./textures/clay-ground-seamless.jpg", "textures/tile1/color.jpg"):
        assert np.array_equal(host.open_image(str(tmp_path / key)), bundled_texture_loader(key))
    for name in ("test1", "test2", "test3"):
        from_disk = host.load_scene(str(tmp_path / "examples" / f"{name}.yml"), texture_root=str(tmp_path))
        assert_same_scene(from_disk, host.parse_scene(example_yaml(name), bundled_texture_loader))
    with pytest.raises(host.HostError) as e:
        host.load_scene(str(tmp_path / "examples" / "test1.yml"), texture_root=str(tmp_path / "nowhere"))
    assert "Could not load texture file ./textures/clay-ground-seamless.jpg" in str(e.value)


def test_gif_decoder_against_pillow(tmp_path):
    """GIF 87a/89a first frame: palettes of several sizes, interlace, transparency, a frame smaller
    than the logical screen.  Compared with Pillow's RGBA view of the same first frame."""
    from PIL import Image

    rng = np.random.default_rng(33)
    base = rng.integers(0, 256, (41, 57, 3), dtype=np.uint8)
    base[8:30, 10:50] = (10, 200, 120)
    rgb = Image.fromarray(base)
    for colors in (2, 4, 16, 256):
        for interlace in (False, True):
            im = rgb.quantize(colors)
            buf = io.BytesIO()
            im.save(buf, "GIF", interlace=interlace)
            data = buf.getvalue()
            mine = host.decode_gif(data)
            want = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))
            assert mine.shape == want.shape and np.array_equal(mine, want), (colors, interlace)
    im = rgb.quantize(16)
    buf = io.BytesIO()
    im.save(buf, "GIF", transparency=3)
    data = buf.getvalue()
    mine = host.decode_gif(data)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))
    assert np.array_equal(mine[..., 3], want[..., 3]) and (mine[..., 3] == 0).any()
    assert np.array_equal(mine[mine[..., 3] == 255], want[want[..., 3] == 255])
    path = tmp_path / "t.gif"
    rgb.quantize(64).save(path, "GIF")
    assert host.open_image(str(path)).shape == (41, 57, 4)
    for bad in (b"", b"GIF89a", data[:30], data[:-40], b"GIF89a" + b"\xff" * 40):
        try:
            host.decode_gif(bad)
        except host.HostError as e:
            assert e.code in (host.E_FORMAT, host.E_UNSUPPORTED, host.E_INVALID)
