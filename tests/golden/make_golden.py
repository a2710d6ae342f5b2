"""Generates tests/golden/*.npz.  diffusers (the module that holds the reference arithmetic) is absent from
/root/reference and not installable, so these fixtures pin (a) closed-form known answers restated in SURVEY.md
Appendix A.8 (sigma tables) and (b) REGRESSION outputs of the oracle itself on the reference's CPU-runnable
configuration (BASELINE.json configs[0]: tiny random-init transformer, 2 blocks, fp32) — they guard the oracle against
accidental edits, they do not prove parity with upstream ("parity unpinned", see oracle header).
Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import qwen_mmdit_ref as R  # noqa: E402

out = Path(__file__).resolve().parent

# (a) sigma tables, SURVEY A.8 (values as printed there, 6 decimals)
np.savez(out / "sigmas_a8.npz",
         n2_s4096=np.array([1.0, 0.02, 0.0]),
         n4_s4096=np.array([1.0, 0.766709, 0.455614, 0.02, 0.0]),
         n8_s4096=np.array([1.0, 0.916024, 0.820046, 0.709295, 0.580075, 0.427347, 0.244054, 0.02, 0.0]),
         n4_s1024=np.array([1.0, 0.749268, 0.432588, 0.02, 0.0]),
         n4_s256=np.array([1.0, 0.744611, 0.426673, 0.02, 0.0]),
         mu_4096=np.array(0.693548), mu_1024=np.array(0.538710), mu_256=np.array(0.5))

# (b) tiny-config oracle outputs (configs[0]): 2 blocks, 4 heads x 32, 256x256 image -> 256 + 256 tokens, T = 19
torch.manual_seed(0)
m = R.init_weights_(R.QwenImageTransformer2DModelRef(R.TINY_CONFIG), seed=0).eval()
shapes = [[(1, 16, 16), (1, 16, 16)]]
hidden, enc = R.make_inputs(R.TINY_CONFIG, shapes, 19, seed=1)
with torch.no_grad():
    v = m(hidden, enc, None, torch.tensor([1.0]), shapes, [19])[0]
    lat = R.ref_run_denoise(m, hidden[:, :256], hidden[:, 256:], enc, shapes, 2)
    lat_cfg = R.ref_run_denoise(m, hidden[:, :256], hidden[:, 256:], enc, shapes, 4, enc[:, :11] * 0.5, 4.0)
img_f, txt_f = m.pos_embed(shapes, [19])
np.savez_compressed(out / "tiny_oracle.npz", velocity=v.numpy()[0, ::16], final_2step=lat.numpy()[0, ::16],
                    final_4step_cfg=lat_cfg.numpy()[0, ::16],
                    rope_img=torch.view_as_real(img_f).numpy()[::37], rope_txt=torch.view_as_real(txt_f).numpy())
print("wrote", [p.name for p in out.glob("*.npz")])
