"""Seed files for tools/fuzz_host.cpp, taken from the test bundle (small JPEG / PNG variants + the example scenes)."""
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image  # noqa: E402

from raingun_b200.examples import _bundle, example_yaml  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "."
os.makedirs(out, exist_ok=True)
b = _bundle()
open(os.path.join(out, "a.jpg"), "wb").write(b["texture/textures/tile1/color.jpg"])
im = Image.open(io.BytesIO(b["texture/./textures/clay-ground-seamless.jpg"])).resize((96, 64))
im.save(os.path.join(out, "prog.jpg"), "JPEG", progressive=True, quality=85)
im.save(os.path.join(out, "base420.jpg"), "JPEG", subsampling=2, quality=85)
im.save(os.path.join(out, "base422.jpg"), "JPEG", subsampling=1, quality=85, restart_marker_blocks=2)
im.convert("L").save(os.path.join(out, "grey.jpg"), "JPEG")
im.save(os.path.join(out, "rgb.png"))
im.convert("P").save(os.path.join(out, "pal.png"))
im.convert("RGBA").save(os.path.join(out, "rgba.png"))
im.convert("1").save(os.path.join(out, "bit.png"))
im.save(os.path.join(out, "rgb.bmp"))
im.convert("P").save(os.path.join(out, "pal.bmp"))
im.convert("RGBA").save(os.path.join(out, "rgba.bmp"))
im.save(os.path.join(out, "rgb.tga"))
im.save(os.path.join(out, "rle.tga"), compression="tga_rle")
im.convert("P").save(os.path.join(out, "pal.tga"))
im.save(os.path.join(out, "rgb.ppm"))
im.convert("1").save(os.path.join(out, "bit.pbm"))
im.quantize(64).save(os.path.join(out, "pal.gif"))
im.quantize(16).save(os.path.join(out, "lace.gif"), interlace=True, transparency=2)
for n in ("test1", "test2", "test3"):
    open(os.path.join(out, n + ".yml"), "w").write(example_yaml(n))
print("seeds written to", out)
