"""Attention micro-benchmark at the step's shape (B=1, 24 heads, 8192+256 tokens) vs torch SDPA."""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
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
H = 24
for img, txt in ((8192, 256), (8192, 219), (2048, 256)):
    s = K.seq(1, img, txt)
    qkv = torch.randn(K.rows(s), 3 * H * 128, device=dev).bfloat16()
    S = img + txt
    flops = 4.0 * S * S * 128 * H
    qkv_b = qkv.clone()                               # bounded-score variants (0x200) take q pre-multiplied by scale * log2(e)
    qkv_b[:, : H * 128] = (qkv_b[:, : H * 128].float() * (0.08838834764831845 * 1.4426950408889634)).bfloat16()
    for v in ([int(x, 0) for x in sys.argv[1:]] or (0x20, 0x100, 0x28, 0x220, 0x230, 0x240, 0x300)):
        ms = bench(lambda: K.attn(s, qkv_b if v & 0x200 else qkv, H, v))
        print(f"attn img={img} txt={txt} variant={v:#x}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    x = qkv[: S].reshape(1, S, 3, H, 128)
    q, k, v_ = (x[:, :, i].transpose(1, 2).contiguous() for i in range(3))
    ms = bench(lambda: F.scaled_dot_product_attention(q, k, v_))
    print(f"torch SDPA S={S}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
