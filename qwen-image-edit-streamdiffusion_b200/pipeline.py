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


def model_timestep(sigma: float, batch: int, device) -> torch.Tensor:
    """Pipeline rounding chain: t = 1000*sigma is cast to the latents dtype (bf16) before /1000."""
    t = torch.tensor([sigma * 1000.0], dtype=torch.float32).to(torch.bfloat16)
    return (t / 1000).expand(batch).to(device)


@torch.no_grad()
def run_denoise(transformer, latents: torch.Tensor, image_latents: torch.Tensor, prompt_embeds: torch.Tensor,
                img_shapes: List, num_inference_steps: int, negative_prompt_embeds: Optional[torch.Tensor] = None,
                true_cfg_scale: float = 4.0, sigmas: Optional[Sequence[float]] = None, collect: Optional[list] = None,
                uncond_fn=None, use_caches: bool = False) -> torch.Tensor:
    """The hot loop.  `uncond_fn(x, ts)` lets the CFG-pair parallel path supply v_uncond from the peer GPU."""
    latents = latents.to(torch.bfloat16).contiguous().clone()
    B, n, _ = latents.shape
    sig = np.asarray(sigmas, dtype=np.float32) if sigmas is not None else flowmatch_sigmas(num_inference_steps, n)
    do_cfg = true_cfg_scale > 1 and (negative_prompt_embeds is not None or uncond_fn is not None)
    image_latents = image_latents.to(torch.bfloat16)
    for i in range(num_inference_steps):
        x = torch.cat([latents, image_latents], dim=1)
        ts_host = model_timestep(float(sig[i]), B, "cpu")
        ts = ts_host.to(latents.device)
        # use_caches: the transformer holds cache_schedule()/cache_prompt("cond"/"uncond") entries (exact)
        kw = dict(timestep_values=[float(ts_host[0])], cached_prompt="cond") if use_caches else {}
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
