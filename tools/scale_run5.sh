#!/bin/bash
# BASELINE configs[4] on one 8-GPU box: streaming batches of 8 frames at 512x512 with two reference images, 4 steps, true CFG.
#   dp8          : 8 frames at once, one per GPU (replicas; every GPU runs the cond and the uncond forward of its frame)
#   cfg+ulysses  : the batch of 8 stays WHOLE inside each of the 2 CFG groups, 4-way fused Ulysses inside the group
mkdir -p gpurun_out
run() {
  name=$1; n=$2; shift 2
  timeout -s KILL 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29547 \
    bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --workload 512x2ref --cfg "$@" > gpurun_out/scale5_r2_$name.json 2> gpurun_out/scale5_r2_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale5_r2_$name.json").read().strip().splitlines()[-1])
    print("$name", "frames/s", round(d["value"], 2), "ms per batch", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 2), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/scale5_r2_$name.err").read()[-1500:])
PY
}
run dp8_b1 8 --mode dp --batch 1
run cfguly8_fused_b8 8 --mode cfg+ulysses --fused --batch 8
