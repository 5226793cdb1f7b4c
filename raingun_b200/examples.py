"""Access to the reference's shipped example scenes, textures and golden renders.

They travel as one bundle, tests/golden/reference_examples.npz (built by
tests/golden/make_fixtures.py from the read-only reference checkout), because
/root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import os
from functools import lru_cache
from typing import Dict

import numpy as np

from .scene import SceneData, parse_scene

_BUNDLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                       "reference_examples.npz")

EXAMPLES = ("test1", "test2", "test3")


@lru_cache(maxsize=1)
def _bundle() -> Dict[str, bytes]:
    with np.load(_BUNDLE) as z:
        return {k: z[k].tobytes() for k in z.files}


@lru_cache(maxsize=8)
def _decode_texture(path: str) -> np.ndarray:
    from . import host

    data = _bundle().get("texture/" + path)
    if data is None:
        raise FileNotFoundError(path)
    # the native decoder restates jpeg-decoder 0.1.11 (host/rgh_jpeg.cpp): with its texels the
    # oracle reproduces examples/test1.png and test3.png bit for bit
    return host.decode_image(data)


def bundled_texture_loader(path: str) -> np.ndarray:
    """Texture loader resolving the path strings used by the example YAMLs."""
    return _decode_texture(path)


def example_yaml(name: str) -> str:
    return _bundle()[f"scene/{name}.yml"].decode("utf-8")


def example_scene(name: str) -> SceneData:
    """examples/<name>.yml as the reference CLI run from its repo root would load it."""
    return parse_scene(example_yaml(name), bundled_texture_loader)


def example_golden(name: str) -> np.ndarray:
    """The committed 800x600 RGBA render examples/<name>.png."""
    from . import host

    img = host.decode_png(_bundle()[f"golden/{name}.png"])
    if img.shape[2] == 3:
        img = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2)
    return img
