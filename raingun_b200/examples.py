"""Access to the reference's shipped example scenes, textures and golden renders.

They travel as one bundle, tests/golden/reference_examples.npz (built by
tests/golden/make_fixtures.py from the read-only reference checkout), because
/root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import io
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
    from PIL import Image

    data = _bundle().get("texture/" + path)
    if data is None:
        raise FileNotFoundError(path)
    with Image.open(io.BytesIO(data)) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8).copy()


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
    from PIL import Image

    with Image.open(io.BytesIO(_bundle()[f"golden/{name}.png"])) as im:
        return np.asarray(im.convert("RGBA"), dtype=np.uint8).copy()
