"""ORACLE — test infrastructure only. Not a product path.

CPU (or eager-GPU) fp32 PyTorch restatement of the one hot path of
shi3z/Qwen-Image-Edit-StreamDiffusion: the `QwenImageTransformer2DModel` denoise step that
every reference script reaches through `pipeline(...)` (reference call sites:
server.py:137-153, qwen_realtime.py:247-255, webui_realtime.py:77-85,
benchmark_lightning.py:24-32; operator swap seam benchmark_lightning_compile.py:89-93).

PARITY UNPINNED.  The arithmetic of this path lives in the un-vendored, un-pinned third-party
dependency `diffusers` (README.md:49, `pip install torch diffusers ...`; first release holding
`QwenImageEditPlusPipeline` is 0.36.0).  diffusers is absent from /root/reference and cannot be
installed here, and the reference's own tests hold no assertion, golden vector or fixture for
this path.  This file therefore restates the *published* algorithm of
diffusers 0.36.0:
    models/transformers/transformer_qwenimage.py   (model, block, attn processor, rope, time embed)
    models/attention_processor.py::Attention       (projection / qk-norm layout)
    models/attention.py::FeedForward, models/activations.py::GELU(approximate="tanh")
    models/normalization.py::{RMSNorm, AdaLayerNormContinuous}
    models/embeddings.py::{get_timestep_embedding, Timesteps, TimestepEmbedding}
    pipelines/qwenimage/pipeline_qwenimage_edit_plus.py::__call__   (true-CFG combine + rescale)
    schedulers/scheduling_flow_match_euler_discrete.py::{set_timesteps, step}
as specified in SURVEY.md Appendix A.1-A.11.  Module / parameter names equal the diffusers
state_dict keys (Appendix A.10) so a real checkpoint or a later diffusers cross-check is one
`load_state_dict` away.  Known answers it is pinned against: the closed-form sigma tables of
Appendix A.8, RoPE invariants and algebraic identities (tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# configuration (Qwen/Qwen-Image-Edit-2509 transformer/config.json; SURVEY Appendix A header)
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RefConfig:
    patch_size: int = 2
    in_channels: int = 64
    out_channels: int = 16
    num_layers: int = 60
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 3584
    guidance_embeds: bool = False
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


TINY_CONFIG = RefConfig(num_layers=2, attention_head_dim=32, num_attention_heads=4,
                        joint_attention_dim=64, axes_dims_rope=(8, 12, 12))
FULL_CONFIG = RefConfig()


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
class RMSNorm(nn.Module):
    """diffusers models/normalization.py::RMSNorm (elementwise_affine=True, no bias)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        in_dtype = x.dtype
        var = x.float().pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.eps)          # promoted to fp32 by the multiply
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(self.weight.dtype)
        x = x * self.weight
        return x.to(in_dtype) if in_dtype == torch.float32 else x


def get_timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool,
                           downscale_freq_shift: float, scale: float,
                           max_period: int = 10000) -> torch.Tensor:
    """diffusers models/embeddings.py::get_timestep_embedding (Appendix A.2)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32,
                                                    device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(F.silu(self.linear_1(sample)))


class QwenTimestepProjEmbeddings(nn.Module):
    """Timesteps(256, flip_sin_to_cos=True, shift=0, scale=1000) + TimestepEmbedding."""

    def __init__(self, embedding_dim: int):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, embedding_dim)

    def forward(self, timestep: torch.Tensor, hidden_states: torch.Tensor) -> torch.Tensor:
        proj = get_timestep_embedding(timestep, 256, flip_sin_to_cos=True,
                                      downscale_freq_shift=0, scale=1000)
        return self.timestep_embedder(proj.to(hidden_states.dtype))


class QwenEmbedRope(nn.Module):
    """transformer_qwenimage.py::QwenEmbedRope(theta, axes_dim, scale_rope=True) (Appendix A.5)."""

    def __init__(self, theta: int, axes_dim: Sequence[int], scale_rope: bool = True):
        super().__init__()
        self.theta = theta
        self.axes_dim = list(axes_dim)
        self.scale_rope = scale_rope
        pos_index = torch.arange(4096)
        neg_index = torch.arange(4096).flip(0) * -1 - 1
        self.pos_freqs = torch.cat([self.rope_params(pos_index, d, theta) for d in self.axes_dim], dim=1)
        self.neg_freqs = torch.cat([self.rope_params(neg_index, d, theta) for d in self.axes_dim], dim=1)

    @staticmethod
    def rope_params(index: torch.Tensor, dim: int, theta: int = 10000) -> torch.Tensor:
        assert dim % 2 == 0
        freqs = torch.outer(index.float(),
                            1.0 / torch.pow(theta, torch.arange(0, dim, 2).float().div(dim)))
        return torch.polar(torch.ones_like(freqs), freqs)

    def forward(self, video_fhw, txt_seq_lens: Sequence[int]):
        if isinstance(video_fhw, list) and len(video_fhw) and isinstance(video_fhw[0], list):
            video_fhw = video_fhw[0]
        if not isinstance(video_fhw, list):
            video_fhw = [video_fhw]
        split = [d // 2 for d in self.axes_dim]
        vid_freqs = []
        max_vid_index = 0
        for idx, (f, h, w) in enumerate(video_fhw):
            fp = self.pos_freqs.split(split, dim=1)
            fn = self.neg_freqs.split(split, dim=1)
            fr = fp[0][idx: idx + f].view(f, 1, 1, -1).expand(f, h, w, -1)
            if self.scale_rope:
                hh = torch.cat([fn[1][-(h - h // 2):], fp[1][: h // 2]], dim=0)
                ww = torch.cat([fn[2][-(w - w // 2):], fp[2][: w // 2]], dim=0)
            else:
                hh, ww = fp[1][:h], fp[2][:w]
            hh = hh.view(1, h, 1, -1).expand(f, h, w, -1)
            ww = ww.view(1, 1, w, -1).expand(f, h, w, -1)
            vid_freqs.append(torch.cat([fr, hh, ww], dim=-1).reshape(f * h * w, -1))
            if self.scale_rope:
                max_vid_index = max(h // 2, w // 2, max_vid_index)
            else:
                max_vid_index = max(h, w, max_vid_index)
        max_len = max(txt_seq_lens)
        txt_freqs = self.pos_freqs[max_vid_index: max_vid_index + max_len]
        return torch.cat(vid_freqs, dim=0), txt_freqs


def apply_rotary_emb_qwen(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """use_real=False branch: adjacent pairs as complex numbers (Appendix A.4)."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    y = torch.view_as_real(xc * freqs.unsqueeze(1)).flatten(3)
    return y.type_as(x)


class _GELUProj(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    """diffusers FeedForward(dim, dim_out=dim, mult=4, activation_fn='gelu-approximate')."""

    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class Attention(nn.Module):
    """Attention(query_dim, added_kv_proj_dim, qk_norm='rms_norm', bias=True) +
    QwenDoubleStreamAttnProcessor2_0 (Appendix A.4)."""

    def __init__(self, dim: int, heads: int, dim_head: int, eps: float = 1e-6):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        inner = heads * dim_head
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.add_q_proj, self.add_k_proj, self.add_v_proj = (nn.Linear(dim, inner), nn.Linear(dim, inner),
                                                             nn.Linear(dim, inner))
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])
        self.to_add_out = nn.Linear(inner, dim)
        self.norm_q, self.norm_k = RMSNorm(dim_head, eps), RMSNorm(dim_head, eps)
        self.norm_added_q, self.norm_added_k = RMSNorm(dim_head, eps), RMSNorm(dim_head, eps)

    def forward(self, img, txt, rotary):
        B, T = txt.shape[0], txt.shape[1]
        H, Dh = self.heads, self.dim_head
        iq, ik, iv = (l(img).unflatten(-1, (H, Dh)) for l in (self.to_q, self.to_k, self.to_v))
        tq, tk, tv = (l(txt).unflatten(-1, (H, Dh)) for l in (self.add_q_proj, self.add_k_proj, self.add_v_proj))
        iq, ik = self.norm_q(iq), self.norm_k(ik)
        tq, tk = self.norm_added_q(tq), self.norm_added_k(tk)
        img_freqs, txt_freqs = rotary
        iq, ik = apply_rotary_emb_qwen(iq, img_freqs), apply_rotary_emb_qwen(ik, img_freqs)
        tq, tk = apply_rotary_emb_qwen(tq, txt_freqs), apply_rotary_emb_qwen(tk, txt_freqs)
        q = torch.cat([tq, iq], dim=1)      # TEXT FIRST
        k = torch.cat([tk, ik], dim=1)
        v = torch.cat([tv, iv], dim=1)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                           attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).flatten(2, 3).to(q.dtype)
        txt_o, img_o = o[:, :T], o[:, T:]
        return self.to_out[0](img_o), self.to_add_out(txt_o)


class QwenImageTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, eps: float = 1e-6):
        super().__init__()
        self.dim = dim
        self.img_mod = nn.Sequential(nn.SiLU(), nn.Linear(dim, 6 * dim))
        self.img_norm1 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.attn = Attention(dim, heads, dim_head, eps)
        self.img_norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.img_mlp = FeedForward(dim)
        self.txt_mod = nn.Sequential(nn.SiLU(), nn.Linear(dim, 6 * dim))
        self.txt_norm1 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.txt_norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.txt_mlp = FeedForward(dim)

    @staticmethod
    def _modulate(x, mod):
        shift, scale, gate = mod.chunk(3, dim=-1)
        return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1), gate.unsqueeze(1)

    def forward(self, h, e, temb, rotary):
        img_mod1, img_mod2 = self.img_mod(temb).chunk(2, dim=-1)
        txt_mod1, txt_mod2 = self.txt_mod(temb).chunk(2, dim=-1)
        img_m, img_g1 = self._modulate(self.img_norm1(h), img_mod1)
        txt_m, txt_g1 = self._modulate(self.txt_norm1(e), txt_mod1)
        img_attn, txt_attn = self.attn(img_m, txt_m, rotary)
        h = h + img_g1 * img_attn
        e = e + txt_g1 * txt_attn
        img_m2, img_g2 = self._modulate(self.img_norm2(h), img_mod2)
        h = h + img_g2 * self.img_mlp(img_m2)
        txt_m2, txt_g2 = self._modulate(self.txt_norm2(e), txt_mod2)
        e = e + txt_g2 * self.txt_mlp(txt_m2)
        return e, h


class AdaLayerNormContinuous(nn.Module):
    """scale FIRST then shift in the chunk (Appendix A.3 note)."""

    def __init__(self, dim: int, cond_dim: int, eps: float = 1e-6):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 2 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)

    def forward(self, x, cond):
        emb = self.linear(F.silu(cond).to(x.dtype))
        scale, shift = emb.chunk(2, dim=1)
        return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


class QwenImageTransformer2DModelRef(nn.Module):
    """Top-level forward, Appendix A.1.  Same keyword surface as the diffusers module."""

    def __init__(self, cfg: RefConfig = FULL_CONFIG):
        super().__init__()
        self.cfg = cfg
        D = cfg.inner_dim
        self.pos_embed = QwenEmbedRope(10000, list(cfg.axes_dims_rope), scale_rope=True)
        self.time_text_embed = QwenTimestepProjEmbeddings(D)
        self.txt_norm = RMSNorm(cfg.joint_attention_dim, eps=1e-6)
        self.img_in = nn.Linear(cfg.in_channels, D)
        self.txt_in = nn.Linear(cfg.joint_attention_dim, D)
        self.transformer_blocks = nn.ModuleList(
            [QwenImageTransformerBlock(D, cfg.num_attention_heads, cfg.attention_head_dim)
             for _ in range(cfg.num_layers)])
        self.norm_out = AdaLayerNormContinuous(D, D, eps=1e-6)
        self.proj_out = nn.Linear(D, cfg.patch_size * cfg.patch_size * cfg.out_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, encoder_hidden_states_mask=None,
                timestep=None, img_shapes=None, txt_seq_lens=None, guidance=None,
                attention_kwargs=None, controlnet_block_samples=None, return_dict=True,
                num_blocks: Optional[int] = None):
        h = self.img_in(hidden_states)
        ts = timestep.to(device=h.device, dtype=h.dtype)
        e = self.txt_in(self.txt_norm(encoder_hidden_states))
        temb = self.time_text_embed(ts, h)
        img_freqs, txt_freqs = self.pos_embed(img_shapes, txt_seq_lens)
        rotary = (img_freqs.to(h.device), txt_freqs.to(h.device))
        blocks = self.transformer_blocks if num_blocks is None else self.transformer_blocks[:num_blocks]
        for blk in blocks:
            e, h = blk(h, e, temb, rotary)
        h = self.norm_out(h, temb)
        out = self.proj_out(h)
        return (out,)


# ----------------------------------------------------------------------------------------------
# random-init recipe: plain N(0,1) explodes through 60 blocks, so fix one recipe.
# ----------------------------------------------------------------------------------------------
def init_weights_(model: nn.Module, seed: int = 0, std: float = 0.02) -> nn.Module:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1 and ("norm" in name) and name.endswith("weight"):
                p.copy_(1.0 + std * torch.randn(p.shape, generator=g))
            elif p.dim() == 2:
                fan_in = p.shape[1]
                # unit-variance-preserving scale for the big matrices, small for modulation
                s = std if ("_mod." in name or "norm_out" in name) else 1.0 / math.sqrt(fan_in)
                p.copy_(s * torch.randn(p.shape, generator=g))
            else:
                p.copy_(std * torch.randn(p.shape, generator=g))
    return model


def make_inputs(cfg: RefConfig, img_shapes, txt_len: int, batch: int = 1, seed: int = 1,
                embed_sigma: float = 3.0, outliers: bool = True):
    """Synthetic latents + cached prompt embeddings of the shapes the pipeline feeds (A.7)."""
    g = torch.Generator().manual_seed(seed)
    s_img = sum(f * h * w for (f, h, w) in img_shapes[0])
    hidden = torch.randn(batch, s_img, cfg.in_channels, generator=g)
    g2 = torch.Generator().manual_seed(seed + 1)
    enc = torch.randn(batch, txt_len, cfg.joint_attention_dim, generator=g2) * embed_sigma
    if outliers:   # Qwen2.5-VL hidden states carry a few massive channels
        idx = torch.randint(0, cfg.joint_attention_dim, (4,), generator=g2)
        enc[..., idx] *= 50.0
    return hidden, enc


# ----------------------------------------------------------------------------------------------
# scheduler + CFG (Appendix A.6 / A.8)
# ----------------------------------------------------------------------------------------------
def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 8192,
                    base_shift: float = 0.5, max_shift: float = 0.9) -> float:
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def ref_flowmatch_sigmas(num_steps: int, image_seq_len: int, shift_terminal: float = 0.02) -> np.ndarray:
    """FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=linspace(1,1/N,N), mu=...) with
    use_dynamic_shifting, time_shift_type='exponential', shift_terminal=0.02; returns N+1 sigmas."""
    sigmas = np.linspace(1.0, 1.0 / num_steps, num_steps).astype(np.float32)
    mu = calculate_shift(image_seq_len)
    sigmas = math.exp(mu) / (math.exp(mu) + (1.0 / sigmas - 1.0) ** 1.0)
    one_minus = 1.0 - sigmas
    scale = one_minus[-1] / (1.0 - shift_terminal)
    sigmas = 1.0 - one_minus / scale
    return np.concatenate([sigmas.astype(np.float32), np.zeros(1, np.float32)])


def ref_cfg_combine(v_cond: torch.Tensor, v_uncond: torch.Tensor, true_cfg_scale: float) -> torch.Tensor:
    comb = v_uncond + true_cfg_scale * (v_cond - v_uncond)
    cond_norm = torch.norm(v_cond, dim=-1, keepdim=True)
    noise_norm = torch.norm(comb, dim=-1, keepdim=True)
    return comb * (cond_norm / noise_norm)


def ref_euler_step(sample: torch.Tensor, model_output: torch.Tensor, sigma: float, sigma_next: float):
    prev = sample.float() + (sigma_next - sigma) * model_output.float()
    return prev.to(model_output.dtype)


def ref_timestep_for_model(sigma: float, dtype: torch.dtype) -> torch.Tensor:
    """Pipeline rounding chain (A.6): t=1000*sigma cast to latents dtype, then /1000."""
    t = torch.tensor([sigma * 1000.0], dtype=torch.float32).to(dtype)
    return t / 1000


def ref_pack_latents(z: torch.Tensor, mean=None, std=None) -> torch.Tensor:
    """QwenImageEditPlusPipeline._pack_latents: [B,C,h,w] -> [B,(h/2)(w/2),4C], channel = c*4 + dy*2 + dx;
    image latents are normalised first: (z - latents_mean) / latents_std (per latent channel)."""
    B, C, h, w = z.shape
    if mean is not None:
        z = (z - mean.view(1, C, 1, 1)) / std.view(1, C, 1, 1)
    return z.view(B, C, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (h // 2) * (w // 2), C * 4)


def ref_unpack_latents(tokens: torch.Tensor, h: int, w: int, mean=None, std=None) -> torch.Tensor:
    """_unpack_latents, then the decode-side de-normalisation z * latents_std + latents_mean (A.7) -> [B,C,1,h,w]."""
    B, n, c4 = tokens.shape
    C = c4 // 4
    z = tokens.view(B, h // 2, w // 2, C, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, C, h, w)
    if mean is not None:
        z = z * std.view(1, C, 1, 1) + mean.view(1, C, 1, 1)
    return z.unsqueeze(2)


def ref_stream_prepare_latent(prev_latent, noise, frame_count: int, keyframe_interval: int = 20, noise_strength: float = 0.05):
    """RealtimeConfig / QwenRealtimePipeline.prepare_latent of the reference (qwen_realtime.py:52-53, 201-224):
    key frame (every keyframe_interval frames, or no previous latent) -> pure noise; else prev_latent + noise_strength*noise."""
    is_key = (frame_count % keyframe_interval == 0) or (prev_latent is None)
    return (noise if is_key else prev_latent + noise_strength * noise), is_key


def ref_run_denoise(model, latents, image_latents, cond_embeds, img_shapes, num_steps: int,
                    uncond_embeds=None, true_cfg_scale: float = 4.0, collect=None, begin_index: int = 0,
                    timestep_dtype: Optional[torch.dtype] = None):
    """The upstream pipeline denoise loop (A.6) on cached embeddings; begin_index = scheduler.set_begin_index.
    `timestep_dtype`: the pipeline casts t = 1000 * sigma to the LATENTS' dtype before dividing by 1000 (A.6), so a bf16
    pipeline feeds the transformer 0.768 where an fp32 one feeds 0.766709; give torch.bfloat16 to follow the bf16 pipeline's
    rounding chain while everything else stays fp32 (default: the latents' own dtype)."""
    sig = ref_flowmatch_sigmas(num_steps, latents.shape[1])
    n = latents.shape[1]
    B = latents.shape[0]
    for i in range(begin_index, num_steps):
        x = torch.cat([latents, image_latents], dim=1)
        ts = ref_timestep_for_model(float(sig[i]), timestep_dtype or latents.dtype).to(latents.dtype).expand(B)
        v = model(hidden_states=x, timestep=ts, encoder_hidden_states=cond_embeds, img_shapes=img_shapes,
                  txt_seq_lens=[cond_embeds.shape[1]] * B, return_dict=False)[0][:, :n]
        if uncond_embeds is not None and true_cfg_scale > 1:
            u = model(hidden_states=x, timestep=ts, encoder_hidden_states=uncond_embeds,
                      img_shapes=img_shapes, txt_seq_lens=[uncond_embeds.shape[1]] * B,
                      return_dict=False)[0][:, :n]
            v = ref_cfg_combine(v, u, true_cfg_scale)
        if collect is not None:
            collect.append(v.clone())
        latents = ref_euler_step(latents, v, float(sig[i]), float(sig[i + 1]))
    return latents


# ----------------------------------------------------------------------------------------------
# W8A8 oracle for the README-only Int8Linear / triton_int8_gemm path (README.md:136-141)
# ----------------------------------------------------------------------------------------------
def ref_int8_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """quantize + matmul + dequantize: symmetric per-out-channel weight, per-token activation."""
    s_w = w.abs().amax(dim=1, keepdim=True).clamp_min(1e-12) / 127.0
    wq = torch.round(w / s_w).clamp(-127, 127)
    s_x = x.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / 127.0
    xq = torch.round(x / s_x).clamp(-127, 127)
    acc = xq.double() @ wq.double().t()          # exact int32 accumulate
    y = acc.float() * s_x * s_w.t()
    return y + bias if bias is not None else y


def ref_fp8_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """e4m3 analogue of ref_int8_linear (per-out-channel weight scale, per-token activation scale)."""
    fmax = 448.0
    s_w = w.abs().amax(dim=1, keepdim=True).clamp_min(1e-12) / fmax
    wq = (w / s_w).to(torch.float8_e4m3fn).float()
    s_x = x.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / fmax
    xq = (x / s_x).to(torch.float8_e4m3fn).float()
    y = (xq @ wq.t()) * s_x * s_w.t()
    return y + bias if bias is not None else y


def ref_llm_int8_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], threshold: float = 6.0) -> torch.Tensor:
    """LLM.int8() as bitsandbytes' Linear8bitLt runs it with `llm_int8_threshold=6.0` — the only int8 configuration that is
    actually present in the reference snapshot (benchmark_int8.py:72-76, test_quantized.py:40-43; the library itself is an
    absent, un-pinned dependency, so this restates its published algorithm, Dettmers et al. 2022 / bitsandbytes
    MatMul8bitLt): feature columns of the activation that hold any |x| > threshold are OUTLIER columns; they are taken out of
    the int8 product (zeroed before the per-token absmax and the rounding) and multiplied in 16 bit against the dequantised
    int8 weight columns; everything else is the vector-wise W8A8 product of ref_int8_linear."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    s_w = w.abs().amax(dim=1, keepdim=True).clamp_min(1e-12) / 127.0
    wq = torch.round(w / s_w).clamp(-127, 127)
    outlier = (x2.abs() > threshold).any(dim=0)                       # [K] feature columns
    x_main = torch.where(outlier[None, :], torch.zeros_like(x2), x2)
    s_x = x_main.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / 127.0
    xq = torch.round(x_main / s_x).clamp(-127, 127)
    y = (xq.double() @ wq.double().t()).float() * s_x * s_w.t()
    if outlier.any():
        w_deq = (wq * s_w)[:, outlier]                                # the stored weights are int8: dequantised columns
        xo = x2[:, outlier].to(torch.bfloat16).float()                # the outlier part runs in 16 bit
        y = y + xo @ w_deq.to(torch.bfloat16).float().t()
    y = y.reshape(*shp[:-1], w.shape[0])
    return y + bias if bias is not None else y


QUANT_LINEARS = {"int8": ref_int8_linear, "fp8": ref_fp8_linear, "llm_int8": ref_llm_int8_linear}
# the four per-block linear groups the W8A8 path replaces (the matrices that hold 99.6 % of the weights and FLOPs): QKV of both
# streams, the two attention output projections, FF up and FF down of both streams.  Embeddings, modulation and proj_out stay 16 bit.
QUANT_TARGETS = ("attn.to_q", "attn.to_k", "attn.to_v", "attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj", "attn.to_out.0",
                 "attn.to_add_out", "img_mlp.net.0.proj", "img_mlp.net.2", "txt_mlp.net.0.proj", "txt_mlp.net.2")


class QuantLinearRef(nn.Module):
    """an nn.Linear evaluated through one of the restated quantised products (same parameters, no copy)"""

    def __init__(self, linear: nn.Linear, kind: str):
        super().__init__()
        self.linear, self.kind = linear, kind

    def forward(self, x):
        y = QUANT_LINEARS[self.kind](x.float(), self.linear.weight.float(), None if self.linear.bias is None else self.linear.bias.float())
        return y.to(x.dtype)


def quantized_view(model: "QwenImageTransformer2DModelRef", kind: str) -> "QwenImageTransformer2DModelRef":
    """The int8 oracle MODEL of config 4: a shallow copy of `model` (parameters shared) whose four per-block linear groups run
    through ref_int8_linear / ref_fp8_linear / ref_llm_int8_linear — 'the reference's own int8 path' restated at model level
    (README.md:136-141 Int8Linear via quantize_transformer.py: replace nn.Linear of the transformer blocks)."""
    import copy
    assert kind in QUANT_LINEARS
    memo = {id(p): p for p in model.parameters()}          # share every parameter tensor
    q = copy.deepcopy(model, memo)
    for blk in q.transformer_blocks:
        for path in QUANT_TARGETS:
            parent = blk
            parts = path.split(".")
            for name in parts[:-1]:
                parent = parent[int(name)] if name.isdigit() else getattr(parent, name)
            leaf = parts[-1]
            lin = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
            wrapped = QuantLinearRef(lin, kind)
            if leaf.isdigit():
                parent[int(leaf)] = wrapped
            else:
                setattr(parent, leaf, wrapped)
    return q.eval()
