"""GEMM micro-benchmark on the step's shapes (through the C ABI), vs torch.matmul (cuBLAS) on the same shapes."""
import math, sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import kernels as K

dev = "cuda:0"
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

import os
K.L.check(K.L.lib().qie_tune(2, int(os.environ.get("QIE_L2_HINTS", "0"))))
s = K.seq(1, 8192, 256)
M = K.rows(s)
D = 3072
res = []
for name, N, Kd, epi in [("qkv", 3 * D, D, K.L.EPI_QKV_NORM_ROPE), ("qkv_plain", 3 * D, D, K.L.EPI_BF16), ("out", D, D, K.L.EPI_GATE_RESID_F32),
                         ("out_plain", D, D, K.L.EPI_BF16), ("ff1", 4 * D, D, K.L.EPI_GELU_BF16), ("ff1_plain", 4 * D, D, K.L.EPI_BF16),
                         ("ff2", D, 4 * D, K.L.EPI_GATE_RESID_F32), ("ff2_plain", D, 4 * D, K.L.EPI_BF16)]:
    a = torch.randn(M, Kd, device=dev).bfloat16()
    w = [(torch.randn(N, Kd, device=dev) / math.sqrt(Kd)).bfloat16() for _ in range(2)]
    b = [torch.randn(N, device=dev) * 0.1 for _ in range(2)]
    f32 = epi in (K.L.EPI_GATE_RESID_F32, K.L.EPI_F32)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    gate = torch.randn(1, 2, 6 * D, device=dev)
    rope = torch.randn(M, 64, 2, device=dev)
    nw = [[torch.ones(128, device=dev) for _ in range(2)] for _ in range(2)]
    flops = 2.0 * M * N * Kd
    for bn, cg in ((256, 2),):
        fn = lambda: K.gemm(s, a, w, b, out, epi, gate=gate, gate_bstride=12 * D, gate_sstride=6 * D, rope=rope, qk_norm_w=nw, block_n=bn, cta_group=cg)
        ms = bench(fn)
        res.append((name, bn, ms, flops / ms / 1e9))
        print(f"{name:10s} bn={bn:3d} cg={cg} {ms:8.3f} ms {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
    ms = bench(lambda: torch.matmul(a, w[0].t()))
    print(f"{name:10s} cuBLAS {ms:8.3f} ms {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
