# Round-end measurement run on one B200: GPU test groups, smoke, the default bench line, the 8-bit modes, the energy probe.
set -x
mkdir -p gpurun_out
TMO=400 bash scripts_gpu_tests.sh > gpurun_out/gpu_tests_summary.log 2>&1
grep -E "exit=|passed|failed" gpurun_out/gpu_tests_summary.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
if [ "$1" != "quick" ]; then timeout 900 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo bench rc=$?; fi
for P in bf16 fp8 int8; do timeout 300 python bench.py --precision $P --no-cpu-baseline --no-strong --no-eager-baseline > gpurun_out/bench_$P.json 2> gpurun_out/bench_$P.err; echo $P rc=$?; done
timeout 300 python tools/energy_probe.py 1.2 > gpurun_out/energy.jsonl 2> gpurun_out/energy.err; echo energy rc=$?
timeout 200 python tools/q8_gemm_bench.py 1.0 > gpurun_out/q8_gemm.jsonl 2> gpurun_out/q8_gemm.err; echo q8 rc=$?
