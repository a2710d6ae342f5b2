# ncu evidence of the current build on one B200 (each capture only after its plain run exited 0; numbers printed under ncu are never
# bench values).  Outputs in gpurun_out/: r02_launches.csv (launch list of a short bench run), r02_gemm / r02_attn / r02_glue .ncu-rep.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-eager-baseline"
# (-c 2400: ~5.5 forwards; ncu serialises and replays, ~0.1 s per launch — the uncapped list of a bench run, ~7800 launches, took more
# than the 8 minutes it was given in round 2)
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 -k regex:"gemm_kernel|attn_|ln_mod|gemv_kernel|cfg_euler|pack_rows|rmsnorm|timestep_proj|quant_rows|unpack_rows" --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_bench.log 2>&1; echo launchlist rc=$?
python tools/ncu_target.py gemm > gpurun_out/plain_gemm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 8 -f -o gpurun_out/r02_gemm python tools/ncu_target.py gemm > gpurun_out/ncu_gemm.log 2>&1; echo gemm rc=$?
python tools/ncu_target.py attn 0x230 0x230 > gpurun_out/plain_attn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_pair -s 1 -c 1 -f -o gpurun_out/r02_attn python tools/ncu_target.py attn 0x230 0x230 > gpurun_out/ncu_attn.log 2>&1; echo attn rc=$?
python tools/ncu_target.py glue > gpurun_out/plain_glue.log 2>&1 && ncu --set full --clock-control none -k regex:"ln_mod_cta|ln_mod_stream" -s 1 -c 1 -f -o gpurun_out/r02_glue python tools/ncu_target.py glue > gpurun_out/ncu_glue.log 2>&1; echo glue rc=$?
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
python tools/ncu_target.py q8 > gpurun_out/plain_q8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -f -o gpurun_out/r02_q8_qkv python tools/ncu_target.py q8 > gpurun_out/ncu_q8.log 2>&1; echo q8 rc=$?
