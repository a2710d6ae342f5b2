#!/bin/bash
# BASELINE configs[4] on one 8-GPU box: 512x512 frames with two reference images, 4 steps, true CFG.
#   dp8         : 8 frames at once, one per GPU (replicas)
#   cfg+ulysses : ONE frame over 2 CFG branches x 4-way Ulysses (fused peer-memory exchange / NCCL)
mkdir -p gpurun_out
run() {
  name=$1; n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29547 \
    bench.py --gpus $n --steps 4 --warmup 3 --no-cpu-baseline --workload 512x2ref --cfg "$@" > gpurun_out/scale5_r1_$name.json 2> gpurun_out/scale5_r1_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale5_r1_$name.json").read().strip().splitlines()[-1])
    print("$name", "frames/s", round(d["value"], 2), "ms_per_step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 2), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/scale5_r1_$name.err").read()[-1500:])
PY
}
run dp8 8 --mode dp
run cfguly8_fused 8 --mode cfg+ulysses --fused
run cfguly8_nccl 8 --mode cfg+ulysses
