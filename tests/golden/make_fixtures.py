"""Builds tests/golden/reference_examples.npz from the read-only reference checkout.

Run HERE (the GPU box has no /root/reference):  python tests/golden/make_fixtures.py

The bundle holds, byte for byte, the reference's only artefacts that pin the render
path: the three example scenes, the JPEG textures they name and
the committed 800x600 renders.  Scenes and textures are INPUT DATA of configs[0..1]
of BASELINE.json, the PNGs are the golden outputs; no reference source code is copied.
"""
import hashlib
import os
import sys

import numpy as np

REF = os.environ.get("RAINGUN_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_examples.npz")

FILES = {
    "scene/test1.yml": "examples/test1.yml",
    "scene/test2.yml": "examples/test2.yml",
    "scene/test3.yml": "examples/test3.yml",
    "golden/test1.png": "examples/test1.png",
    "golden/test2.png": "examples/test2.png",
    "golden/test3.png": "examples/test3.png",
    # keyed by the path strings the YAML files use (relative to the reference root)
    "texture/./textures/clay-ground-seamless.jpg": "textures/clay-ground-seamless.jpg",
    "texture/./textures/land_ocean_ice_cloud_2048.jpg": "textures/land_ocean_ice_cloud_2048.jpg",
    "texture/textures/tile1/color.jpg": "textures/tile1/color.jpg",
}


def main() -> int:
    blobs = {}
    for key, rel in FILES.items():
        with open(os.path.join(REF, rel), "rb") as f:
            data = f.read()
        blobs[key] = np.frombuffer(data, dtype=np.uint8)
        print(f"{key:55s} {len(data):8d} B sha256={hashlib.sha256(data).hexdigest()[:16]}")
    np.savez(OUT, **blobs)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
