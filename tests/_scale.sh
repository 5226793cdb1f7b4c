#!/bin/bash
# scaling sweep on one box: bash tests/_scale.sh "1 2 4 8"
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.1f Mrays/s  %.2f ms/frame | e2e %.1f Mrays/s %.2f ms (upload %.2f) | launches %d | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["scene_upload_ms_per_step"], d["gpu_launches"], d["config"]["parallelism"]))
    il = d.get("in_library_multi_gpu")
    if il: print("      in-library: static %.2f ms/frame, steal %.2f ms/frame" % (il["static"]["ms_per_frame"], il["steal"]["ms_per_frame"]))
    c5 = (d.get("configs") or {}).get("C5@7680x4320")
    if c5: print("      C5: 1 GPU %.1f ms" % c5["ms_per_frame"], ("| sharded x%d %.1f ms" % (c5["sharded"]["n_gpus"], c5["sharded"]["ms_per_frame_wall_incl_d2h"])) if "sharded" in c5 else "")
except Exception as e:
    print("N=$n ERR", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
# optional extra: tile heights at the largest N (bash tests/_scale.sh "8" "4 16")
for tr in ${2:-}; do
  n=$(echo $1 | awk '{print $NF}')
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+tr)) bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline --no-roofline --no-configs --no-in-library --tile-rows $tr > gpurun_out/scale_${n}_t$tr.json 2> gpurun_out/scale_${n}_t$tr.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_${n}_t$tr.json").read().strip().splitlines()[-1])
print("N=$n tile-rows $tr: value %.1f Mrays/s %.2f ms | e2e %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
done
