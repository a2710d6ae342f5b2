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
    outs = {}
    # 2 = CTA-row form (modulation vectors in registers; library default), 1 = warp-per-row streaming ring, 0 = one warp per row
    for v in (1, 2, 0):
        K.L.check(K.L.lib().qie_tune(3, v))
        ms = bench(run)
        outs[v] = K.ln_modulate(s, xs[0], mod, 12 * D, 6 * D, 0, D, D).float()
        print(f"B={B} ln variant {v}: {ms*1e3:.1f} us {byts/ms/1e6:.0f} GB/s", flush=True)
        if v != 0:
            for qm, name in ((1, "e4m3"), (2, "int8")):
                ms = bench(lambda: K.ln_modulate(s, xs[0], mod, 12 * D, 6 * D, 0, D, D, fp8=True, qmode=qm))
                print(f"B={B} ln variant {v} + {name} shadow: {ms*1e3:.1f} us", flush=True)
                ms = bench(lambda: K.ln_modulate(s, xs[0], mod, 12 * D, 6 * D, 0, D, D, fp8=True, qmode=qm, want_bf16=False))
                print(f"B={B} ln variant {v}, {name} rows only (the W8A8 forward): {ms*1e3:.1f} us", flush=True)
    print(f"B={B} variant 2 vs 1: max |diff| {(outs[2] - outs[1]).abs().max().item():.3e} "
          f"(bf16 ulp at max |y| = {outs[1].abs().max().item() * 2 ** -8:.3e}), rows differing "
          f"{((outs[2] != outs[1]).any(dim=1)).sum().item()} of {outs[1].shape[0]}", flush=True)
    K.L.check(K.L.lib().qie_tune(3, 2))
    ms = bench(cp)
    print(f"B={B} torch fp32->bf16 copy: {ms*1e3:.1f} us {byts/ms/1e6:.0f} GB/s", flush=True)
