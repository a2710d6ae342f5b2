import os, sys
sys.path.insert(0,"tests"); sys.path.insert(0,".")
import kernels as K
for mode in (0, 1, 9):
    K.L.check(K.L.lib().qie_tune(4, mode))
    print("== split mode", mode, flush=True)
    src = open("tools/gemm_bench.py").read().replace('("qkv", 3 * D, D, K.L.EPI_QKV_NORM_ROPE), ', '').replace('("ff1", 4 * D, D, K.L.EPI_GELU_BF16), ("ff1_plain", 4 * D, D, K.L.EPI_BF16),', '')
    src = src.replace("    ms = bench(lambda: torch.matmul(a, w[0].t()))\n", "    continue\n    ms = 0\n")
    exec(src)
