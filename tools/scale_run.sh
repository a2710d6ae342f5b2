#!/bin/bash
# Multi-GPU bench matrix on one 8-GPU box: Ulysses / CFG-pair x Ulysses, NCCL all-to-all vs fused peer-memory exchange.
# usage: tools/scale_run.sh  (writes gpurun_out/scale_r1_*.json)
mkdir -p gpurun_out
run() {  # name nproc devices args...
  name=$1; n=$2; devs=$3; shift 3
  CUDA_VISIBLE_DEVICES=$devs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port 29541 bench.py --gpus $n --steps 4 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_r1_$name.json 2> gpurun_out/scale_r1_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_r1_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms_per_image", round(d["ms_per_step"], 1), "e2e_ms", round(d["e2e"]["ms_per_step"], 1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/scale_r1_$name.err").read()[-1500:])
PY
}
run uly8_fused 8 0,1,2,3,4,5,6,7 --mode ulysses --fused
run cfguly8_fused 8 0,1,2,3,4,5,6,7 --mode cfg+ulysses --cfg --fused
run uly8_nccl 8 0,1,2,3,4,5,6,7 --mode ulysses
run cfguly8_nccl 8 0,1,2,3,4,5,6,7 --mode cfg+ulysses --cfg
run uly4_fused 4 0,1,2,3 --mode ulysses --fused
run cfguly4_fused 4 0,1,2,3 --mode cfg+ulysses --cfg --fused
