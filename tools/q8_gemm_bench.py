"""8-bit GEMM probe: the four per-block linears of config 4 in e4m3 (kind::f8f6f4) and int8 (kind::i8) through the C ABI, with
their production epilogues, against the library on the same box — cuBLASLt e4m3 (`torch._scaled_mm`, per-tensor scales, bf16 out)
and cuBLASLt int8 (`torch._int_mm`) — burst (best of 10 after warm-up) and sustained (back to back for SUSTAIN_S seconds, the
figure that applies inside a power-capped step).  Prints one JSON line per (shape, kind): the measured 8-bit peaks the W8A8 path is
held against ("no measured FP8 peak to put 2027 TFLOP/s against").

usage: python tools/q8_gemm_bench.py [SUSTAIN_S=1.5]"""
import json, math, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import kernels as K

dev = "cuda:0"
SUSTAIN_S = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5


def burst(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def sustained(fn, seconds):
    """back-to-back launches for `seconds`; the mean of the second half (clocks settled under the power cap)"""
    fn(); torch.cuda.synchronize()
    t_end = time.time() + seconds
    ms, chunk = [], 20
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(chunk): fn()
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / chunk)
    half = ms[len(ms) // 2:]
    return sum(half) / len(half)


s = K.seq(1, 8192, 256)
M = K.rows(s)
D = 3072
one = torch.ones((), device=dev)
for name, N, Kd, epi in [("qkv+rmsnorm+rope", 3 * D, D, K.L.EPI_QKV_NORM_ROPE), ("out-proj+gate+resid", D, D, K.L.EPI_GATE_RESID_F32),
                         ("ff-up+gelu", 4 * D, D, K.L.EPI_GELU_BF16), ("ff-down+gate+resid", D, 4 * D, K.L.EPI_GATE_RESID_F32)]:
    flops = 2.0 * M * N * Kd
    f32 = epi == K.L.EPI_GATE_RESID_F32
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    b = [torch.randn(N, device=dev) * 0.1 for _ in range(2)]
    gate = torch.randn(1, 2, 6 * D, device=dev)
    rope = torch.randn(M, 64, 2, device=dev)
    nw = [[torch.ones(128, device=dev) for _ in range(2)] for _ in range(2)]
    a_sc = torch.full((M,), 0.01, device=dev)
    w_sc = [torch.full((N,), 0.01, device=dev) for _ in range(2)]
    for kind, mode in (("fp8", 1), ("int8", 2)):
        if kind == "fp8":
            a8 = torch.randn(M, Kd, device=dev).to(torch.float8_e4m3fn)
            w8 = [(torch.randn(N, Kd, device=dev)).to(torch.float8_e4m3fn) for _ in range(2)]
            lib_fn = lambda: torch._scaled_mm(a8, w8[0].t(), scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
        else:
            a8 = torch.randint(-127, 128, (M, Kd), device=dev, dtype=torch.int8)
            w8 = [torch.randint(-127, 128, (N, Kd), device=dev, dtype=torch.int8) for _ in range(2)]
            wt = w8[0].t()
            lib_fn = lambda: torch._int_mm(a8, wt)
        au, wu = a8.view(torch.uint8), [w.view(torch.uint8) for w in w8]
        ours = lambda: K.gemm(s, au, wu, b, out, epi, gate=gate, gate_bstride=12 * D, gate_sstride=6 * D, rope=rope, qk_norm_w=nw,
                              fp8=mode, a_scale=a_sc, w_scale=w_sc, block_n=256, cta_group=2)
        rec = {"shape": name, "M": M, "N": N, "K": Kd, "kind": kind}
        for tag, fn in (("qie", ours), ("cublaslt", lib_fn)):
            try:
                rec[tag + "_burst_tops"] = round(flops / burst(fn) / 1e9, 1)
                rec[tag + "_sustained_tops"] = round(flops / sustained(fn, SUSTAIN_S) / 1e9, 1)
            except Exception as e:       # a library entry point this torch build does not offer on sm_100: say so, keep going
                rec[tag + "_error"] = str(e)[:160]
        if "qie_sustained_tops" in rec and "cublaslt_sustained_tops" in rec:
            rec["ratio_sustained"] = round(rec["qie_sustained_tops"] / rec["cublaslt_sustained_tops"], 3)
        print(json.dumps(rec), flush=True)
