import sys
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
s = K.seq(1, 8192, 256); D = 3072
xs = [torch.randn(K.rows(s), D, device=dev) for _ in range(4)]   # rotate buffers > L2
mod = torch.randn(1, 2, 6 * D, device=dev)
i = [0]
def run():
    i[0] = (i[0] + 1) % 4
    K.ln_modulate(s, xs[i[0]], mod, 12 * D, 6 * D, 0, D, D)
byts = K.rows(s) * D * 6
K.L.check(K.L.lib().qie_tune(3, 1))
ms = bench(run)
print(f"streaming kernel: {ms*1e3:.1f} us {byts/ms/1e6:.0f} GB/s", flush=True)
K.L.check(K.L.lib().qie_tune(3, 0))
for thr in (128,):
    for smem in (0,):
        K.L.check(K.L.lib().qie_tune(0, thr)); K.L.check(K.L.lib().qie_tune(1, smem * 1024))
        ms = bench(run)
        print(f"threads {thr} smem {smem}KB: {ms*1e3:.1f} us {byts/ms/1e6:.0f} GB/s", flush=True)
