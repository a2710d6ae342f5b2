"""Tail handling of the persistent GEMM schedule on the step's shapes and on the token shards of the sequence-parallel forward
(qie_tune key 4): 0 = whole tiles, 1 = K split of long-K tails, 17 = + N split where no K split applies (default), 49 = N split
everywhere.  usage: python tools/gemm_tail_probe.py"""
import math, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import kernels as K

dev = "cuda:0"
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

D = 3072
for label, img, txt in (("1 GPU", 8192, 256), ("2-way shard", 4096, 128), ("4-way shard", 2048, 64), ("8-way shard", 1024, 32)):
    s = K.seq(1, img, txt)
    M = K.rows(s)
    for name, N, Kd, epi in [("qkv", 3 * D, D, K.L.EPI_QKV_NORM_ROPE), ("out", D, D, K.L.EPI_GATE_RESID_F32),
                             ("ff1", 4 * D, D, K.L.EPI_GELU_BF16), ("ff2", D, 4 * D, K.L.EPI_GATE_RESID_F32)]:
        a = torch.randn(M, Kd, device=dev).bfloat16()
        w = [(torch.randn(N, Kd, device=dev) / math.sqrt(Kd)).bfloat16() for _ in range(2)]
        b = [torch.randn(N, device=dev) * 0.1 for _ in range(2)]
        f32 = epi == K.L.EPI_GATE_RESID_F32
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        gate = torch.randn(1, 2, 6 * D, device=dev)
        rope = torch.randn(M, 64, 2, device=dev)
        nw = [[torch.ones(128, device=dev) for _ in range(2)] for _ in range(2)]
        row = []
        for mode in (0, 1, 17, 49):
            K.L.check(K.L.lib().qie_tune(4, mode))
            row.append(bench(lambda: K.gemm(s, a, w, b, out, epi, gate=gate, gate_bstride=12 * D, gate_sstride=6 * D, rope=rope, qk_norm_w=nw)))
        K.L.check(K.L.lib().qie_tune(4, 17))
        flops = 2.0 * (img + txt) * N * Kd
        print(f"{label:12s} {name:4s} whole {row[0] * 1e3:7.1f} us | K split {row[1] * 1e3:7.1f} | default {row[2] * 1e3:7.1f} ({flops / row[2] / 1e9:6.0f} TFLOP/s) | N split everywhere {row[3] * 1e3:7.1f}", flush=True)
