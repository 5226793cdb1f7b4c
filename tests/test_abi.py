"""The C-ABI library loads without a GPU and exports every symbol include/raingun_b200.h
declares; argument validation that needs no device works; rendering without a device fails
loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from raingun_b200 import _native
from raingun_b200.examples import example_scene
from raingun_b200.scene import SceneDesc, Stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "raingun_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z_0-9]+)\s*\(", text)) - {"rg_rows_cb"})


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    # offsets implied by include/raingun_b200.h on LP64
    assert SceneDesc.fov.offset == 8 and SceneDesc.default_color.offset == 16
    assert SceneDesc.n_bodies.offset == 28 and SceneDesc.body_kind.offset == 32
    assert SceneDesc.n_lights.offset == 104 and SceneDesc.light_kind.offset == 112
    assert SceneDesc.textures.offset == 144 and ctypes.sizeof(SceneDesc) == 152
    assert ctypes.sizeof(Stats) == 184


def test_validation_without_device():
    lib = _native.lib()
    out = ctypes.c_void_p()
    assert lib.rg_scene_create(None, 0, ctypes.byref(out)) == _native.E_INVALID
    assert b"NULL" in lib.rg_last_error()
    data = example_scene("test2")
    data.max_recursion_depth = 65
    desc, keep = data.to_desc()
    assert lib.rg_scene_create(ctypes.byref(desc), 0, ctypes.byref(out)) == _native.E_DEPTH
    data.max_recursion_depth = 10
    data.body_kind = data.body_kind.copy()
    data.body_kind[0] = 9
    desc, keep = data.to_desc()
    assert lib.rg_scene_create(ctypes.byref(desc), 0, ctypes.byref(out)) == _native.E_INVALID
    assert lib.rg_render(None, 4, 4, None, None) == _native.E_INVALID


def test_no_cpu_fallback():
    """Without a usable GPU scene creation must fail with RG_E_CUDA, never render on the CPU."""
    lib = _native.lib()
    if lib.rg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    desc, keep = example_scene("test2").to_desc()
    out = ctypes.c_void_p()
    assert lib.rg_scene_create(ctypes.byref(desc), 0, ctypes.byref(out)) == _native.E_CUDA
    assert b"no CPU path" in lib.rg_last_error()
    a = ctypes.c_double()
    assert lib.rg_measure_peaks(0, ctypes.byref(a), ctypes.byref(a), ctypes.byref(a)) == _native.E_CUDA


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under raingun_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "raingun_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "raingun_oracle" not in text and "from oracle" not in text and "import oracle" not in text, \
                    os.path.join(dirpath, f)


def test_headers_are_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: both headers must compile as C99 (no C++-isms, no torch or
    CUDA types in the signatures) and a C program must link against the libraries' entry points."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "raingun_b200.h"\n#include "raingun_host.h"\n'
        "int main(void) {\n"
        "    rg_scene_desc d; rg_stats s; rgh_image im; rgh_cli_options o; rg_scene *sc = 0; rgh_scene *hs = 0;\n"
        "    (void)d; (void)s; (void)im; (void)o;\n"
        "    /* take the address of every entry point a host would bind */\n"
        "    void *fns[] = {(void *)rg_scene_create, (void *)rg_scene_create_multi, (void *)rg_scene_device_count, (void *)rg_scene_destroy, (void *)rg_scene_set_option, (void *)rg_render,\n"
        "                   (void *)rg_render_rows, (void *)rg_render_rows_device, (void *)rg_render_rowlist_device,\n"
        "                   (void *)rg_render_rowlist_scatter, (void *)rg_render_rowlist_host, (void *)rg_host_register, (void *)rg_host_unregister, (void *)rg_device_enable_peer, (void *)rg_shm_barrier_open, (void *)rg_shm_barrier_wait, (void *)rg_shm_barrier_close, (void *)rg_shared_frame_create, (void *)rg_shared_frame_open,\n"
        "                   (void *)rg_shared_frame_close, (void *)rg_render_stream, (void *)rg_render_rows_f32, (void *)rg_render_stream_f32, (void *)rg_trim, (void *)rg_last_error, (void *)rg_device_count,\n"
        "                   (void *)rgh_scene_load, (void *)rgh_scene_parse, (void *)rgh_scene_desc, (void *)rgh_scene_destroy,\n"
        "                   (void *)rgh_jpeg_decode, (void *)rgh_png_decode, (void *)rgh_png_encode, (void *)rgh_image_open,\n"
        "                   (void *)rgh_png_save, (void *)rgh_cli_parse, (void *)rgh_last_error, (void *)rgh_free};\n"
        "    (void)sc; (void)hs;\n"
        "    return rg_device_count() < 0 || sizeof(fns) == 0;\n"
        "}\n")
    pkg = os.path.join(ROOT, "raingun_b200")
    exe = tmp_path / "abi"
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-Wno-pedantic", "-I", inc, str(src), "-o", str(exe),
                        "-L", pkg, "-lraingun_b200", "-lraingun_host", f"-Wl,-rpath,{pkg}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr   # rg_device_count() is 0 without a GPU, never negative


def test_rust_sys_crate_mirrors_the_header():
    """integration/raingun-b200-sys cannot be compiled here (no cargo), so at least keep its struct
    fields and prototypes in step with include/raingun_b200.h: same names, same order."""
    hdr = open(os.path.join(ROOT, "include", "raingun_b200.h")).read()
    rs = open(os.path.join(ROOT, "integration", "raingun-b200-sys", "src", "lib.rs")).read()

    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return [re.sub(r"\[.*", "", d.strip().split()[-1].lstrip("*")) for d in body.split(";") if d.strip()]

    def rs_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, re.S).group(1)
        return re.findall(r"pub (\w+):", body)

    for name in ("rg_texture_desc", "rg_scene_desc", "rg_stats"):
        assert c_fields(name) == rs_fields(name), name
    c_fns = set(re.findall(r"\b(rg_[a-z_0-9]+)\s*\(", hdr)) - {"rg_rows_cb", "rg_rows_f32_cb"}
    rs_fns = set(re.findall(r"pub fn (rg_[a-z_0-9]+)\(", rs))
    assert c_fns == rs_fns == set(_native.EXPORTS)


def test_c_example_host_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """integration/example_host.c: the whole drop-in from plain C.  Here (no GPU) it must load the
    scene natively and then stop at the upload with the library's message - never render on the CPU."""
    import shutil
    import subprocess

    from raingun_b200 import device_count
    from raingun_b200.examples import example_yaml

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    pkg = os.path.join(ROOT, "raingun_b200")
    exe = tmp_path / "example_host"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "example_host.c"), "-o", str(exe), "-L", pkg,
                        "-lraingun_host", "-lraingun_b200", f"-Wl,-rpath,{pkg}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    scene = tmp_path / "test2.yml"
    scene.write_text(example_yaml("test2"))
    assert subprocess.run([str(exe)], capture_output=True).returncode == 2
    r = subprocess.run([str(exe), str(tmp_path / "missing.yml"), str(tmp_path / "o.png")], capture_output=True, text=True)
    assert r.returncode == 3 and "Could not open input file" in r.stderr
    r = subprocess.run([str(exe), str(scene), str(tmp_path / "o.png"), "160", "120"], capture_output=True, text=True)
    if device_count() == 0:
        assert r.returncode == 4 and "no CUDA device" in r.stderr and not (tmp_path / "o.png").exists()
    else:
        assert r.returncode == 0 and (tmp_path / "o.png").exists()
