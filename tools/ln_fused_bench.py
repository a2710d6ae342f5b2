"""GATE_RESID GEMM + following adaLN: separate launches vs adaLN fused into the GEMM tail (standalone, config-2 shapes)."""
import sys, math, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import kernels as K
DEV = "cuda:0"
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
import os
K.L.check(K.L.lib().qie_tune(6, int(os.environ.get("QIE_LN_DBG", "0"))))
s = K.seq(1, 8192, 256); rows = K.rows(s); N = 3072
gate = torch.randn(1, 2, 6 * N, device=DEV)
for name, Kd in (("out", 3072), ("ff2", 12288)):
    a = torch.randn(rows, Kd, device=DEV).bfloat16()
    w = [(torch.randn(N, Kd, device=DEV) / math.sqrt(Kd)).bfloat16() for _ in range(2)]
    b = [torch.randn(N, device=DEV) * 0.1 for _ in range(2)]
    res = torch.randn(rows, N, device=DEV)
    xm = torch.empty(rows, N, dtype=torch.bfloat16, device=DEV)
    ln = dict(out=xm, mod=gate, bstride=12 * N, sstride=6 * N, shift_off=3 * N, scale_off=4 * N)
    g = lambda l=None: K.gemm(s, a, w, b, res, K.L.EPI_GATE_RESID_F32, gate=gate[:, :, 2 * N:], gate_bstride=12 * N, gate_sstride=6 * N, cta_group=2, ln=l)
    t_g = bench(lambda: g())
    t_sep = bench(lambda: (g(), K.ln_modulate(s, res, gate, 12 * N, 6 * N, 3 * N, 4 * N, N)))
    t_f = bench(lambda: g(ln))
    print(f"{name}: gemm {t_g*1e3:.1f} us | gemm + ln kernel {t_sep*1e3:.1f} us | fused {t_f*1e3:.1f} us", flush=True)
