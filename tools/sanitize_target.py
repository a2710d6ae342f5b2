"""Small end-to-end run for compute-sanitizer (memcheck): tiny 2-block model, ragged shapes, every kernel class once."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import qie_b200
cfg = qie_b200.QwenImageDiTConfig(num_layers=2, num_attention_heads=2, joint_attention_dim=128)
m = qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=0, device="cuda:0")
shapes = [[(1, 10, 12), (1, 16, 8)]]
g = torch.Generator(device="cuda:0").manual_seed(0)
lat = torch.randn(1, 120, 64, generator=g, device="cuda:0").bfloat16()
img = torch.randn(1, 128, 64, generator=g, device="cuda:0").bfloat16()
c = torch.randn(1, 37, 128, generator=g, device="cuda:0").bfloat16()
u = torch.randn(1, 20, 128, generator=g, device="cuda:0").bfloat16()
for prec in ("bf16", "fp8", "int8"):
    m.set_precision(prec)
    out = qie_b200.run_denoise(m, lat, img, c, shapes, 2, u, 4.0)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
