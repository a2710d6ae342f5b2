"""Denoise loop around the transformer: the part of QwenImageEditPlusPipeline.__call__ that sits on the hot path
(SURVEY A.6): per step one (cond) or two (cond + uncond, true-CFG) transformer forwards, then the fused
CFG-combine + norm-rescale + FlowMatch-Euler kernel.  VAE and the Qwen2.5-VL encoder stay on the reference; their
outputs arrive here as cached tensors (image_latents, prompt_embeds), as in the reference's cached_pipeline_v2.py
(README.md:125).  Reference call sites: server.py:137-153, qwen_realtime.py:247-255, webui_realtime.py:77-85.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def flowmatch_sigmas(num_steps: int, image_seq_len: int) -> np.ndarray:
    """FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=linspace(1,1/N,N), mu=calculate_shift(seq)) -> N+1 sigmas."""
    if num_steps < 2:
        # upstream's stretch-to-terminal computes 1 - (1 - sigma) / ((1 - sigma[-1]) / 0.98): with the single sigma 1.0 that is
        # 0 / 0 -> NaN timesteps (and NaN images) without any error; fail loudly instead
        raise L.QieError("the dynamic-shift schedule needs at least 2 steps (one step is 0/0 in the reference scheduler); "
                         "pass explicit `sigmas` to run_denoise for a single-step update")
    buf = (C.c_float * (num_steps + 1))()
    L.check(L.lib().qie_flowmatch_sigmas(num_steps, image_seq_len, buf), "qie_flowmatch_sigmas")
    return np.array(buf[:], dtype=np.float32)


def cfg_euler_step(latents: torch.Tensor, v_cond: torch.Tensor, v_uncond: Optional[torch.Tensor],
                   true_cfg_scale: float, sigma: float, sigma_next: float) -> torch.Tensor:
    """In-place: latents <- latents + (sigma_next - sigma) * cfg(v_cond, v_uncond).  bf16 tensors.
    v_* may hold more tokens than latents (the transformer returns noise + reference tokens; only the
    first latents.shape[1] are consumed, exactly like `noise_pred[:, :latents.size(1)]` upstream)."""
    assert latents.dtype == torch.bfloat16 and v_cond.dtype == torch.bfloat16 and latents.is_contiguous()
    assert v_cond.is_contiguous() and (v_uncond is None or (v_uncond.is_contiguous() and v_uncond.shape == v_cond.shape))
    B, n, ch = latents.shape
    with torch.cuda.device(latents.device):
        L.check(L.lib().qie_cfg_euler_step(L.ptr(v_cond), L.ptr(v_uncond), L.ptr(latents), float(true_cfg_scale),
                                           float(sigma), float(sigma_next), B, n, ch, v_cond.shape[1], L.cur_stream()),
                "qie_cfg_euler_step")
    return latents


_TS_CACHE: dict = {}


def model_timestep(sigma: float, batch: int, device) -> torch.Tensor:
    """Pipeline rounding chain: t = 1000*sigma is cast to the latents dtype (bf16) before /1000.
    The device copy of a (sigma, batch, device) triple is made once (a fixed schedule re-uses 2-8 values for every frame): the
    per-step pageable host-to-device copy of one scalar would otherwise block the host once per denoise step."""
    dev = torch.device(device)
    key = (float(sigma), int(batch), dev.type, dev.index)
    t = _TS_CACHE.get(key)
    if t is None:
        h = torch.tensor([sigma * 1000.0], dtype=torch.float32).to(torch.bfloat16)
        t = (h / 1000).expand(batch).contiguous().to(dev)
        if len(_TS_CACHE) > 256:
            _TS_CACHE.clear()
        _TS_CACHE[key] = t
    return t


def pack_latents(z: torch.Tensor, mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`_pack_latents` of the upstream pipeline: [B, C, h, w] (or [B, C, 1, h, w]) bf16 -> [B, (h/2)(w/2), 4C];
    with the VAE `latents_mean` / `latents_std` given, the `(z - mean) / std` of the image-latent path is fused in."""
    if z.dim() == 5:
        z = z[:, :, 0]
    assert z.dtype == torch.bfloat16 and z.is_cuda
    z = z.contiguous()
    B, Cc, h, w = z.shape
    out = torch.empty(B, (h // 2) * (w // 2), 4 * Cc, dtype=torch.bfloat16, device=z.device)
    m = None if mean is None else mean.to(z.device, torch.float32).contiguous()
    sd = None if std is None else std.to(z.device, torch.float32).contiguous()
    with torch.cuda.device(z.device):
        L.check(L.lib().qie_pack_latents(L.ptr(z), L.ptr(m), L.ptr(sd), L.ptr(out), B, Cc, h, w, L.cur_stream()), "qie_pack_latents")
    return out


def unpack_latents(tokens: torch.Tensor, h: int, w: int, mean: Optional[torch.Tensor] = None,
                   std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`_unpack_latents` + `z * latents_std + latents_mean` (decode side): [B, (h/2)(w/2), 4C] -> [B, C, 1, h, w]."""
    assert tokens.dtype == torch.bfloat16 and tokens.is_cuda
    tokens = tokens.contiguous()
    B, n, c4 = tokens.shape
    if n != (h // 2) * (w // 2) or c4 % 4:
        raise L.QieError(f"unpack_latents: {n} tokens x {c4} channels do not form a {h}x{w} latent")
    z = torch.empty(B, c4 // 4, h, w, dtype=torch.bfloat16, device=tokens.device)
    m = None if mean is None else mean.to(tokens.device, torch.float32).contiguous()
    sd = None if std is None else std.to(tokens.device, torch.float32).contiguous()
    with torch.cuda.device(tokens.device):
        L.check(L.lib().qie_unpack_latents(L.ptr(tokens), L.ptr(m), L.ptr(sd), L.ptr(z), B, c4 // 4, h, w, L.cur_stream()),
                "qie_unpack_latents")
    return z.unsqueeze(2)


@torch.no_grad()
def run_denoise(transformer, latents: torch.Tensor, image_latents: torch.Tensor, prompt_embeds: torch.Tensor,
                img_shapes: List, num_inference_steps: int, negative_prompt_embeds: Optional[torch.Tensor] = None,
                true_cfg_scale: float = 4.0, sigmas: Optional[Sequence[float]] = None, collect: Optional[list] = None,
                uncond_fn=None, use_caches: bool = False, begin_index: int = 0, batched_cfg: bool = False) -> torch.Tensor:
    """The hot loop.  `uncond_fn(x, ts)` lets the CFG-pair parallel path supply v_uncond from the peer GPU.
    `begin_index` > 0 enters the schedule part-way (scheduler.set_begin_index upstream): the streaming path starts from the
    previous frame's re-noised latent instead of pure noise.
    `batched_cfg`: the cond and the uncond forward of a step run as ONE forward of batch 2B (the reference's
    batched_cfg_pipeline.py, README.md:126), each half with its own text length — worth it when one frame does not fill the GPU
    (512x512 frames); the shorter prompt's rows are padding, not attended tokens, so the velocities equal the separate forwards."""
    latents = latents.to(torch.bfloat16).contiguous().clone()
    B, n, _ = latents.shape
    sig = np.asarray(sigmas, dtype=np.float32) if sigmas is not None else flowmatch_sigmas(num_inference_steps, n)
    do_cfg = true_cfg_scale > 1 and (negative_prompt_embeds is not None or uncond_fn is not None)
    image_latents = image_latents.to(torch.bfloat16)
    batched = bool(batched_cfg and do_cfg and uncond_fn is None and not use_caches and 2 * B <= 8)
    if batched:
        Tc, Tu = prompt_embeds.shape[1], negative_prompt_embeds.shape[1]
        both = torch.zeros(2 * B, max(Tc, Tu), prompt_embeds.shape[2], dtype=torch.bfloat16, device=latents.device)
        both[:B, :Tc] = prompt_embeds
        both[B:, :Tu] = negative_prompt_embeds
        shapes2 = (list(img_shapes) * 2 if isinstance(img_shapes[0][0], (list, tuple)) and len(img_shapes) == B else img_shapes)
    for i in range(begin_index, num_inference_steps):
        x = torch.cat([latents, image_latents], dim=1)
        ts = model_timestep(float(sig[i]), B, latents.device)
        if batched:
            vu = transformer(hidden_states=torch.cat([x, x], 0), timestep=torch.cat([ts, ts], 0), encoder_hidden_states=both,
                             img_shapes=shapes2, txt_seq_lens=[Tc] * B + [Tu] * B, return_dict=False)[0]
            v, u = vu[:B].contiguous(), vu[B:].contiguous()
            if collect is not None:
                collect.append((v[:, :n].clone(), u[:, :n].clone()))
            cfg_euler_step(latents, v, u, true_cfg_scale, float(sig[i]), float(sig[i + 1]))
            continue
        # use_caches: the transformer holds cache_schedule()/cache_prompt("cond"/"uncond") entries (exact)
        kw = dict(timestep_values=[float(model_timestep(float(sig[i]), 1, "cpu")[0])], cached_prompt="cond") if use_caches else {}
        v = transformer(hidden_states=x, timestep=ts, encoder_hidden_states=prompt_embeds, img_shapes=img_shapes,
                        txt_seq_lens=[prompt_embeds.shape[1]] * B, return_dict=False, **kw)[0]
        u = None
        if do_cfg:
            if uncond_fn is not None:
                u = uncond_fn(x, ts)
            else:
                kwu = dict(kw, cached_prompt="uncond") if use_caches else {}
                u = transformer(hidden_states=x, timestep=ts, encoder_hidden_states=negative_prompt_embeds,
                                img_shapes=img_shapes, txt_seq_lens=[negative_prompt_embeds.shape[1]] * B,
                                return_dict=False, **kwu)[0]
        if collect is not None:
            collect.append((v[:, :n].clone(), None if u is None else u[:, :n].clone()))
        cfg_euler_step(latents, v, u, true_cfg_scale, float(sig[i]), float(sig[i + 1]))
    return latents


# ------------------------------------------------------------------------------------------------------------------
# "next" row N4: stateful streaming (the StreamDiffusion-style loop qwen_realtime.py:201-224 sketches but never wires:
# `prepare_latent` is dead code there and process_frame always denoises from pure noise, qwen_realtime.py:226-268)
# ------------------------------------------------------------------------------------------------------------------
class StreamingDenoiser:
    """Per-frame state of a realtime edit stream.

    Key frames (every `keyframe_interval` frames, and the first) denoise from pure noise over the whole schedule; the frames
    in between start from `prev_latent + noise_strength * noise` (qwen_realtime.py:216-222, `noise_strength` 0.05,
    `keyframe_interval` 20 in RealtimeConfig) and only run the last `stream_steps` steps of the schedule, i.e. they enter
    FlowMatchEulerDiscreteScheduler at `begin_index = num_inference_steps - stream_steps`."""

    def __init__(self, transformer, img_shapes, prompt_embeds, negative_prompt_embeds=None, num_inference_steps: int = 4,
                 true_cfg_scale: float = 1.0, noise_strength: float = 0.05, keyframe_interval: int = 20,
                 stream_steps: int = 2, denoise_fn=None):
        if not (1 <= stream_steps <= num_inference_steps):
            raise ValueError("stream_steps must be in [1, num_inference_steps]")
        if keyframe_interval < 1:
            raise ValueError("keyframe_interval must be >= 1")
        self.t, self.img_shapes = transformer, img_shapes
        self.cond, self.uncond = prompt_embeds, negative_prompt_embeds
        self.steps, self.cfg_scale = num_inference_steps, true_cfg_scale
        self.noise_strength, self.keyframe_interval, self.stream_steps = noise_strength, keyframe_interval, stream_steps
        self.prev_latent: Optional[torch.Tensor] = None
        self.frame_count = 0
        self.is_keyframe = True
        self._denoise = denoise_fn or run_denoise

    def prepare_latent(self, noise: torch.Tensor):
        """qwen_realtime.py:201-224 — returns (start latent, begin_index)."""
        self.is_keyframe = (self.frame_count % self.keyframe_interval == 0) or (self.prev_latent is None)
        if self.is_keyframe:
            return noise, 0
        return self.prev_latent + self.noise_strength * noise.to(self.prev_latent.dtype), self.steps - self.stream_steps

    def process_frame(self, image_latents: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """One frame: `image_latents` = packed VAE latents of the camera frame (condition), `noise` = N(0,1) of the latent shape."""
        start, begin = self.prepare_latent(noise)
        out = self._denoise(self.t, start, image_latents, self.cond, self.img_shapes, self.steps, self.uncond, self.cfg_scale,
                            begin_index=begin)
        self.prev_latent = out
        self.frame_count += 1
        return out

    def forwards_per_frame(self) -> int:
        per_step = 2 if (self.uncond is not None and self.cfg_scale > 1) else 1
        return per_step * (self.steps if self.is_keyframe else self.stream_steps)
