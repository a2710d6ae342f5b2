"""In-step A/B of the launch / kernel knobs: one full-size model, the headline step (2 forwards at 8192+256 tokens) timed
under each setting with CUDA events, baseline re-measured between the groups (the step runs at the power cap, so standalone
gains do not carry over one-to-one; this is the measurement that decides defaults).

    python tools/instep_sweep.py [steps]      ->  JSON lines {"knob", "value", "ms_per_image", "vs_base"}
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import qie_b200  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
L = qie_b200.lib()
cfg = qie_b200.QwenImageDiTConfig()
model = qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=0, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
lat = torch.randn(1, 4096, 64, generator=g, device=dev).bfloat16()
img = torch.randn(1, 4096, 64, generator=g, device=dev).bfloat16()
cond = (torch.randn(1, 256, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16()
shapes = [[(1, 64, 64), (1, 64, 64)]]


def measure():
    for _ in range(2):
        qie_b200.run_denoise(model, lat, img, cond, shapes, 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        qie_b200.run_denoise(model, lat, img, cond, shapes, 2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def tune(key):
    return lambda v: L.qie_tune(key, v)


def option(key):
    return lambda v: model.set_option(key, v)


# (name, setter, default, values)
KNOBS = [
    ("bounded-score attention where the norm weights allow it (set_option 3)", option(3), 1, [0]),
    ("polynomial share of the exponentials (set_option 1; the bounded form adds 0x200)", option(1), 0, [0x100, 0x20, 0x40]),
    ("programmatic dependent launch (tune 7)", tune(7), 1, [0]),
    ("GEMM raster band, m-units (tune 5)", tune(5), 0, [8, 16, 32]),
    ("GEMM TMA L2 hints (tune 2)", tune(2), 0, [1, 3]),
    ("GEMM tail: K split / N split (tune 4; default 17)", tune(4), 17, [0, 1, 16]),
    ("adaLN kernel: 2 = CTA-row form (default), 1 = warp-per-row ring, 0 = one warp per row (tune 3)", tune(3), 2, [1, 0, 1]),
]
if len(sys.argv) > 2:                     # python tools/instep_sweep.py 4 0,1 -> only these knobs
    KNOBS = [KNOBS[int(i)] for i in sys.argv[2].split(",")]

for _ in range(3):          # reach the power-capped steady state before the first number
    measure()
base = measure()
print(json.dumps({"knob": "baseline", "value": None, "ms_per_image": base}), flush=True)
for name, setter, default, values in KNOBS:
    for v in values:
        setter(v)
        ms = measure()
        setter(default)
        print(json.dumps({"knob": name, "value": hex(v) if v > 64 else v, "ms_per_image": ms, "vs_base": base / ms}), flush=True)
    base = measure()
    print(json.dumps({"knob": "baseline", "value": None, "ms_per_image": base}), flush=True)
