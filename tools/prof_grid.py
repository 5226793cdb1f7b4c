import sys
sys.path.insert(0, ".")
import raingun_b200 as rg
from raingun_b200.synth import make_scene
sd, spec = make_scene("C4")
sc = rg.Scene(sd); sc.set_accel(2)
for it in range(2):
    sc.render_rows(spec.width, spec.height, 0, spec.height)
print("trace ms", sc.last_stats.ms_trace)
