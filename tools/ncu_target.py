"""Minimal launch sequence for ncu captures: python tools/ncu_target.py attn|gemm|glue|q8 [fp8|int8]"""
import sys
from pathlib import Path
import math, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import kernels as K
import os
if os.environ.get("QIE_GROUP_M"):
    K.L.check(K.L.lib().qie_tune(5, int(os.environ["QIE_GROUP_M"])))
if os.environ.get("QIE_L2_HINTS"):
    K.L.check(K.L.lib().qie_tune(2, int(os.environ["QIE_L2_HINTS"])))
dev = "cuda:0"
what = sys.argv[1]
s = K.seq(1, 8192, 256)
if what == "attn":
    H = 24
    qkv = torch.randn(K.rows(s), 3 * H * 128, device=dev).bfloat16()
    for v in ([int(x, 0) for x in sys.argv[2:]] or [1, 0, 1, 0]):
        K.attn(s, qkv, H, v)
elif what == "glue":
    D = 3072
    x = torch.randn(K.rows(s), D, device=dev)
    mod = torch.randn(1, 2, 6 * D, device=dev)
    for _ in range(3):
        K.ln_modulate(s, x, mod, 12 * D, 6 * D, 0, D, D)
    temb = torch.randn(1, D, device=dev)
    wmod = torch.randn(60 * 2 * 6 * D, D, device=dev).bfloat16()       # all 120 modulation matrices stacked (13.6 GB)
    bmod = torch.randn(60 * 2 * 6 * D, device=dev)
    for _ in range(2):
        K.gemv(temb, wmod, bmod, 1)
elif what == "q8":       # the 8-bit forms of the QKV and out-proj GEMMs (two launches each)
    D = 3072
    M = K.rows(s)
    mode = 2 if (len(sys.argv) > 2 and sys.argv[2] == "int8") else 1
    for name, N, Kd, epi in [("qkv", 3 * D, D, K.L.EPI_QKV_NORM_ROPE), ("out", D, D, K.L.EPI_GATE_RESID_F32)]:
        if mode == 1:
            a8 = torch.randn(M, Kd, device=dev).to(torch.float8_e4m3fn).view(torch.uint8)
            w8 = [torch.randn(N, Kd, device=dev).to(torch.float8_e4m3fn).view(torch.uint8) for _ in range(2)]
        else:
            a8 = torch.randint(-127, 128, (M, Kd), device=dev, dtype=torch.int8).view(torch.uint8)
            w8 = [torch.randint(-127, 128, (N, Kd), device=dev, dtype=torch.int8).view(torch.uint8) for _ in range(2)]
        b = [torch.randn(N, device=dev) * 0.1 for _ in range(2)]
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == K.L.EPI_GATE_RESID_F32 else torch.bfloat16)
        gate = torch.randn(1, 2, 6 * D, device=dev)
        rope = torch.randn(M, 64, 2, device=dev)
        nw = [[torch.ones(128, device=dev) for _ in range(2)] for _ in range(2)]
        a_sc = torch.full((M,), 0.01, device=dev)
        w_sc = [torch.full((N,), 0.01, device=dev) for _ in range(2)]
        for _ in range(2):
            K.gemm(s, a8, w8, b, out, epi, gate=gate, gate_bstride=12 * D, gate_sstride=6 * D, rope=rope, qk_norm_w=nw,
                   fp8=mode, a_scale=a_sc, w_scale=w_sc, block_n=256, cta_group=2)
else:
    D = 3072
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
        for _ in range(2):
            K.gemm(s, a, w, b, out, epi, gate=gate, gate_bstride=12 * D, gate_sstride=6 * D, rope=rope, qk_norm_w=nw)
torch.cuda.synchronize()
