import sys
from pathlib import Path
import torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
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
for B in (1, 4):
    s = K.seq(B, 8192, 256)
    xs = [torch.randn(K.rows(s), D, device=dev) for _ in range(4)]
    ys = [torch.empty(K.rows(s), D, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    mod = torch.randn(B, 2, 6 * D, device=dev)
    i = [0]
    def run():
        i[0] = (i[0] + 1) % 4
        K.ln_modulate(s, xs[i[0]], mod, 12 * D, 6 * D, 0, D, D)
    def cp():
        i[0] = (i[0] + 1) % 4
        ys[i[0]].copy_(xs[i[0]])
    byts = K.rows(s) * D * 6
    for v in (1, 0):
        K.L.check(K.L.lib().qie_tune(3, v))
        ms = bench(run)
        print(f"B={B} ln variant {v}: {ms*1e3:.1f} us {byts/ms/1e6:.0f} GB/s", flush=True)
    ms = bench(cp)
    print(f"B={B} torch fp32->bf16 copy: {ms*1e3:.1f} us {byts/ms/1e6:.0f} GB/s", flush=True)
