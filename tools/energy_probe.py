"""Joules per launch of the step's kernels, each looped alone for ~1.5 s at the power cap.

The single-GPU step is energy-bound, so the number that predicts in-step speed is energy per FLOP / per byte,
not standalone time.  NVML's total-energy counter (mJ) is read before and after a back-to-back loop of one kernel; the
library baselines (cuBLAS through torch.matmul, cuDNN/flash SDPA) run the same shapes for comparison.

    python tools/energy_probe.py [seconds_per_case]   ->  JSON lines
"""
import json
import math
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import kernels as K  # noqa: E402

import pynvml  # noqa: E402  (nvidia_ml_py)

SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
dev = "cuda:0"
pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)


def energy_mj():
    return pynvml.nvmlDeviceGetTotalEnergyConsumption(nv)


def probe(name, fn, work, unit):
    """work = algorithmic FLOPs (unit 'flop') or bytes (unit 'byte') per launch"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(10, int(SECONDS * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
    for _ in range(n // 2):          # reach the sustained clock before the measured loop
        fn()
    torch.cuda.synchronize()
    mj0, t0 = energy_mj(), time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    mj1, t1 = energy_mj(), time.perf_counter()
    ms = e0.elapsed_time(e1) / n
    joule = (mj1 - mj0) * 1e-3 / n
    clock = pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
    rec = {"kernel": name, "launches": n, "ms": round(ms, 4), "joule_per_launch": round(joule, 4),
           "avg_watt": round((mj1 - mj0) * 1e-3 / (t1 - t0), 1), "sm_mhz_after": clock}
    if unit == "flop":
        rec["tflops"] = round(work / ms / 1e9, 1)
        rec["pj_per_flop"] = round(joule / work * 1e12, 4)
    else:
        rec["gbs"] = round(work / ms / 1e6, 1)
        rec["pj_per_byte"] = round(joule / work * 1e12, 3)
    print(json.dumps(rec), flush=True)


s = K.seq(1, 8192, 256)
M, D, H = K.rows(s), 3072, 24
gate = torch.randn(1, 2, 6 * D, device=dev)
rope = torch.randn(M, 64, 2, device=dev)
nw = [[torch.ones(128, device=dev) for _ in range(2)] for _ in range(2)]
for name, N, Kd, epi in [("qkv", 3 * D, D, K.L.EPI_QKV_NORM_ROPE), ("out", D, D, K.L.EPI_GATE_RESID_F32),
                         ("ff1", 4 * D, D, K.L.EPI_GELU_BF16), ("ff2", D, 4 * D, K.L.EPI_GATE_RESID_F32)]:
    a = torch.randn(M, Kd, device=dev).bfloat16()
    w = [(torch.randn(N, Kd, device=dev) / math.sqrt(Kd)).bfloat16() for _ in range(2)]
    b = [torch.randn(N, device=dev) * 0.1 for _ in range(2)]
    f32 = epi == K.L.EPI_GATE_RESID_F32
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    flops = 2.0 * M * N * Kd
    probe(f"gemm_kernel {name} [{M}x{N}x{Kd}]",
          lambda: K.gemm(s, a, w, b, out, epi, gate=gate, gate_bstride=12 * D, gate_sstride=6 * D, rope=rope, qk_norm_w=nw), flops, "flop")
    probe(f"cuBLAS {name} [{M}x{N}x{Kd}] (torch.matmul, no epilogue)", lambda: torch.matmul(a, w[0].t()), flops, "flop")
    del a, w, b, out

qkv = torch.randn(M, 3 * H * 128, device=dev).bfloat16()
S = 8192 + 256
aflops = 4.0 * S * S * 128 * H
qkv_b = qkv.clone()             # the bounded-score form (0x200) takes q pre-multiplied by softmax_scale * log2(e)
qkv_b[:, : H * 128] = (qkv_b[:, : H * 128].float() * (0.08838834764831845 * 1.4426950408889634)).bfloat16()
for v in (0x20, 0x220, 0x230, 0x240):
    probe(f"attn_pair_kernel variant {v:#x}", lambda: K.attn(s, qkv_b if v & 0x200 else qkv, H, v), aflops, "flop")
x = qkv[:S].reshape(1, S, 3, H, 128)
q, k, v_ = (x[:, :, i].transpose(1, 2).contiguous() for i in range(3))
probe("torch SDPA (cuDNN / flash)", lambda: F.scaled_dot_product_attention(q, k, v_), aflops, "flop")

xr = torch.randn(M, D, device=dev)
mod = torch.randn(1, 2, 6 * D, device=dev)
probe("ln_mod_stream_kernel", lambda: K.ln_modulate(s, xr, mod, 12 * D, 6 * D, 0, D, D), (8192 + 256) * D * 6.0, "byte")
for qm, nm in ((1, "e4m3"), (2, "int8")):
    probe(f"ln_mod_stream_kernel + fused {nm} quantiser", lambda: K.ln_modulate(s, xr, mod, 12 * D, 6 * D, 0, D, D, fp8=True, qmode=qm),
          (8192 + 256) * D * 7.0, "byte")
