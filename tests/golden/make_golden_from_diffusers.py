"""Pins the oracle to the REAL upstream implementation whenever `diffusers` (>= 0.36.0, the un-vendored dependency that holds
the arithmetic of this path — reference README.md:49) can be imported.  It cannot in the build image (no network, package not
in the wheelhouse), so nothing here runs there; the file exists so that the first environment that does have diffusers turns
"parity unpinned" into a measured statement:

    python tests/golden/make_golden_from_diffusers.py        # writes tests/golden/diffusers_tiny.npz

The fixture holds, for BASELINE.json configs[0] (tiny random-init QwenImageTransformer2DModel: 2 blocks, 4 heads x 32, rope axes
(8, 12, 12), joint dim 64; 256x256 image -> 256 noise + 256 reference tokens, T = 19, fp32 on CPU, seed 0):
  * the state_dict of the diffusers module (so the oracle is checked on IDENTICAL weights through load_state_dict — the oracle
    keeps the diffusers parameter names),
  * the inputs, the single-step velocity of the diffusers module, its RoPE tables,
  * the sigma / timestep tables of FlowMatchEulerDiscreteScheduler (Qwen-Image scheduler_config) for 2, 4 and 8 steps at
    image_seq_len 4096 / 1024 / 256, and one scheduler.step() result.
tests/test_oracle_cpu.py::test_oracle_matches_diffusers_* compare the oracle with it (auto-skipped while the fixture and
diffusers are both absent); with diffusers importable the same tests ALSO run the comparison live, without the fixture.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

TINY = dict(patch_size=2, in_channels=64, out_channels=16, num_layers=2, attention_head_dim=32, num_attention_heads=4,
            joint_attention_dim=64, guidance_embeds=False, axes_dims_rope=(8, 12, 12))
# scheduler_config.json of Qwen/Qwen-Image-Edit-2509
SCHED = dict(base_image_seq_len=256, max_image_seq_len=8192, base_shift=0.5, max_shift=0.9, shift=1.0, shift_terminal=0.02,
             use_dynamic_shifting=True, time_shift_type="exponential", num_train_timesteps=1000, invert_sigmas=False,
             stochastic_sampling=False, use_beta_sigmas=False, use_exponential_sigmas=False, use_karras_sigmas=False)
SHAPES = [[(1, 16, 16), (1, 16, 16)]]
T_TXT = 19


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=8192, base_shift=0.5, max_shift=0.9):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    return image_seq_len * m + (base_shift - m * base_seq_len)


def build_diffusers_tiny(seed: int = 0):
    """the upstream module, random-init with the recipe of SURVEY 8c (so 2 blocks do not blow up), fp32, eval"""
    from diffusers import QwenImageTransformer2DModel
    torch.manual_seed(seed)
    m = QwenImageTransformer2DModel(**TINY).eval()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("weight") and p.dim() == 1:          # RMSNorm weights
                p.copy_(1.0 + 0.02 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return m


def diffusers_outputs(m, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(1, 512, TINY["in_channels"], generator=g)
    enc = torch.randn(1, T_TXT, TINY["joint_attention_dim"], generator=g) * 3
    ts = torch.tensor([1.0])
    with torch.no_grad():
        kw = dict(hidden_states=hidden, encoder_hidden_states=enc, encoder_hidden_states_mask=torch.ones(1, T_TXT, dtype=torch.long),
                  timestep=ts, img_shapes=SHAPES, return_dict=False)
        try:
            v = m(txt_seq_lens=[T_TXT], **kw)[0]
        except TypeError:                    # later releases dropped txt_seq_lens in favour of the mask
            v = m(**kw)[0]
        try:
            rope = m.pos_embed(SHAPES, [T_TXT], device=hidden.device)
        except TypeError:
            rope = m.pos_embed(SHAPES, max_txt_seq_len=T_TXT, device=hidden.device)
    return hidden, enc, ts, v, rope


def scheduler_tables():
    from diffusers import FlowMatchEulerDiscreteScheduler
    out = {}
    for n in (2, 4, 8):
        for seq in (4096, 1024, 256):
            sch = FlowMatchEulerDiscreteScheduler(**SCHED)
            sch.set_timesteps(sigmas=np.linspace(1.0, 1.0 / n, n), mu=calculate_shift(seq), device="cpu")
            out[f"sigmas_n{n}_s{seq}"] = sch.sigmas.numpy()
            out[f"timesteps_n{n}_s{seq}"] = sch.timesteps.numpy()
    sch = FlowMatchEulerDiscreteScheduler(**SCHED)
    sch.set_timesteps(sigmas=np.linspace(1.0, 0.25, 4), mu=calculate_shift(4096), device="cpu")
    g = torch.Generator().manual_seed(7)
    x, v = torch.randn(1, 64, 64, generator=g), torch.randn(1, 64, 64, generator=g)
    out["step_x"], out["step_v"] = x.numpy(), v.numpy()
    out["step_out"] = sch.step(v, sch.timesteps[0], x, return_dict=False)[0].numpy()
    return out


def main():
    import diffusers
    m = build_diffusers_tiny()
    hidden, enc, ts, v, rope = diffusers_outputs(m)
    blob = {"diffusers_version": np.array(diffusers.__version__), "hidden": hidden.numpy(), "enc": enc.numpy(), "timestep": ts.numpy(),
            "velocity": v.numpy(), "rope_img": torch.view_as_real(rope[0]).numpy(), "rope_txt": torch.view_as_real(rope[1]).numpy()}
    for k, t in m.state_dict().items():
        blob["sd." + k] = t.numpy()
    blob.update(scheduler_tables())
    out = Path(__file__).resolve().parent / "diffusers_tiny.npz"
    np.savez_compressed(out, **blob)
    print("wrote", out, f"(diffusers {diffusers.__version__})")


if __name__ == "__main__":
    main()
