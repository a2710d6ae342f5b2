"""Multi-GPU partitions of the denoise step (one process per GPU, torch.distributed / NCCL over NVLink).

  * CFG pair      — cond / uncond forwards of a step are independent given the same latents: branch 0 runs the cond
                    forward, branch 1 the uncond forward, one tiny all-gather of the velocities (512 KB at 1024^2) and both
                    run the fused CFG+Euler kernel redundantly so latents stay replicated.  Replaces the reference's
                    parallel_cfg_pipeline.py "2 GPUs ... with CUDA streams" (README.md:127-128; source absent).
  * Ulysses SP    — inside a branch the joint sequence is sharded over P ranks for every per-token op (adaLN, QKV, out-proj,
                    FF: replicated weights, 1/P of the tokens) and re-sharded over heads for attention with two
                    all-to-alls per block (24 heads -> 24/P per rank).  The reference has no counterpart.
  * replicas      — independent frames per GPU need no code here (bench.py default at N > 1).

The host-side planning (`make_shard_plan`, `pack_heads` / `unpack_heads`) is pure torch and is covered by world_size-2
gloo tests on CPU; the kernels run through the phase API of libqie.so (qie_forward_phase / qie_attn_fwd_tiles).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


# ------------------------------------------------------------------------------------------------
# host-side planning
# ------------------------------------------------------------------------------------------------
def _pad128(n: int) -> int:
    return (n + 127) // 128 * 128


@dataclass(frozen=True)
class ShardPlan:
    size: int
    rank: int
    img_total: int
    txt_total: int
    img_offset: int
    img_rows: int
    txt_offset: int
    txt_rows: int
    img_pad: int          # identical on every rank
    txt_pad: int          # identical on every rank
    tile_valid: tuple     # valid rows of every 128-row tile of the gathered (rank-major) sequence

    @property
    def rows_pad(self) -> int:
        return self.img_pad + self.txt_pad


def split_sizes(total: int, parts: int) -> List[int]:
    return [total // parts + (1 if r < total % parts else 0) for r in range(parts)]


def make_shard_plan(img_total: int, txt_total: int, size: int, rank: int) -> ShardPlan:
    """Token shard of `rank`: image and text tokens are each split contiguously over the ranks."""
    if size < 1 or not (0 <= rank < size):
        raise ValueError("bad rank/size")
    if img_total < size or txt_total < size:
        raise ValueError(f"every rank needs at least one image and one text token (img={img_total}, txt={txt_total}, P={size})")
    im, tx = split_sizes(img_total, size), split_sizes(txt_total, size)
    img_pad, txt_pad = _pad128(max(im)), _pad128(max(tx))
    if _pad128(min(im)) != img_pad or _pad128(min(tx)) != txt_pad:
        raise ValueError("shards straddle a 128-row boundary: a rank would own an all-padding tile")
    tiles = []
    for r in range(size):
        for n, pad in ((im[r], img_pad), (tx[r], txt_pad)):
            tiles += [min(128, n - t * 128) for t in range(pad // 128)]
    return ShardPlan(size, rank, img_total, txt_total, sum(im[:rank]), im[rank], sum(tx[:rank]), tx[rank], img_pad, txt_pad,
                     tuple(tiles))


def pack_heads(qkv_local: torch.Tensor, size: int, heads: int) -> torch.Tensor:
    """[rows, 3*H*128] (q|k|v, head-major) -> [P, rows, 3*(H/P)*128]: slab g holds heads [g*H/P, (g+1)*H/P) of q, k, v."""
    rows = qkv_local.shape[0]
    hl = heads // size
    return qkv_local.view(rows, 3, size, hl * 128).permute(2, 0, 1, 3).contiguous().view(size, rows, 3 * hl * 128)


def unpack_heads(o_recv: torch.Tensor, size: int) -> torch.Tensor:
    """[P (head group), rows, (H/P)*128] -> [rows, H*128]"""
    _, rows, w = o_recv.shape
    return o_recv.permute(1, 0, 2).reshape(rows, size * w)


# ------------------------------------------------------------------------------------------------
# Ulysses sequence-parallel transformer
# ------------------------------------------------------------------------------------------------
class UlyssesTransformer:
    """Wraps a B200QwenImageTransformer2DModel; same call surface, the work of ONE forward is spread over `group`."""

    def __init__(self, transformer, group=None):
        self.t = transformer
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if transformer.cfg.num_attention_heads % self.size:
            raise L.QieError(f"{transformer.cfg.num_attention_heads} heads do not split over {self.size} ranks")
        self.config = transformer.config
        self._tile_cache = {}

    def cache_context(self, name):
        return self.t.cache_context(name)

    def __call__(self, hidden_states, encoder_hidden_states=None, encoder_hidden_states_mask=None, timestep=None,
                 img_shapes=None, txt_seq_lens=None, guidance=None, attention_kwargs=None, return_dict=True, **_):
        t, lib = self.t, L.lib()
        B, S_i, _ = hidden_states.shape
        if B != 1:
            raise L.QieError("sequence parallelism is implemented for batch 1 (run frames as replicas / CFG branches)")
        T = encoder_hidden_states.shape[1]
        plan = make_shard_plan(S_i, T, self.size, self.rank)
        dev = t.device
        seq = L.Seq(1, plan.img_rows, plan.txt_rows, plan.img_pad, plan.txt_pad)
        sp = L.Sp(plan.rank, plan.size, plan.img_total, plan.txt_total, plan.img_offset, plan.txt_offset)
        ws = t._workspace(seq)
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        ws_bytes = ws.numel() - (base - ws.data_ptr())
        D, H, P = t.cfg.inner_dim, t.cfg.num_attention_heads, self.size
        rows = plan.rows_pad

        def view(which, width):
            off = lib.qie_workspace_offset(t._handle, C.byref(seq), which) + (base - ws.data_ptr())
            return ws[off: off + rows * width * 2].view(torch.bfloat16).view(rows, width)

        qkv_local, attn_local = view(0, 3 * D), view(1, D)
        hs = hidden_states[:, plan.img_offset: plan.img_offset + plan.img_rows].to(torch.bfloat16).contiguous()
        enc = encoder_hidden_states[:, plan.txt_offset: plan.txt_offset + plan.txt_rows].to(dev, torch.bfloat16).contiguous()
        ts = timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(1).contiguous()
        shapes = img_shapes[0] if isinstance(img_shapes[0][0], (list, tuple)) else img_shapes
        flat = [int(v) for fhw in shapes for v in fhw]
        shp = (C.c_int * len(flat))(*flat)
        out_local = torch.empty(1, plan.img_rows, t.cfg.out_dim, dtype=torch.bfloat16, device=dev)
        key = plan.tile_valid
        if key not in self._tile_cache:
            self._tile_cache[key] = torch.tensor(plan.tile_valid, dtype=torch.int32, device=dev)
        tiles = self._tile_cache[key]
        hl = H // P
        q_recv = torch.empty(P, rows, 3 * hl * 128, dtype=torch.bfloat16, device=dev)
        o_full = torch.empty(P, rows, hl * 128, dtype=torch.bfloat16, device=dev)
        o_recv = torch.empty_like(o_full)

        def phase(mask, layer):
            L.check(lib.qie_forward_phase(t._handle, mask, layer, L.ptr(hs), L.ptr(enc), L.ptr(ts), shp, len(flat) // 3,
                                          C.byref(seq), C.byref(sp), L.ptr(out_local), C.c_void_p(base), ws_bytes, -1,
                                          L.cur_stream()), "qie_forward_phase")

        with torch.cuda.device(dev):
            phase(1, -1)
            for l in range(t.cfg.num_layers):
                phase(2, l)                                                          # adaLN1 + QKV on local tokens
                dist.all_to_all_single(q_recv, pack_heads(qkv_local, P, H), group=self.group)   # tokens -> heads
                L.check(lib.qie_attn_fwd_tiles(L.ptr(q_recv), L.ptr(o_full), P * rows // 128, L.ptr(tiles), hl, 0,
                                               L.cur_stream()), "qie_attn_fwd_tiles")
                dist.all_to_all_single(o_recv, o_full, group=self.group)             # heads -> tokens
                attn_local.copy_(unpack_heads(o_recv, P))
                phase(8, l)                                                          # out-proj, adaLN2, FF on local tokens
            phase(16, -1)
            # every rank needs the whole velocity for the (replicated) Euler update
            sizes = split_sizes(S_i, P)
            if len(set(sizes)) == 1:
                full = torch.empty(P * plan.img_rows, t.cfg.out_dim, dtype=torch.bfloat16, device=dev)
                dist.all_gather_into_tensor(full, out_local[0], group=self.group)
                out = full.view(1, S_i, t.cfg.out_dim)
            else:
                parts = [torch.empty(n, t.cfg.out_dim, dtype=torch.bfloat16, device=dev) for n in sizes]
                dist.all_gather(parts, out_local[0].contiguous(), group=self.group)
                out = torch.cat(parts, 0).unsqueeze(0)
        out = out.to(hidden_states.dtype) if hidden_states.dtype != torch.bfloat16 else out
        return (out,) if not return_dict else type("Out", (), {"sample": out})()


# ------------------------------------------------------------------------------------------------
# CFG pair (optionally composed with Ulysses inside each branch)
# ------------------------------------------------------------------------------------------------
@dataclass
class ParallelLayout:
    """world = cfg_branches x sp_size ranks; rank = branch * sp_size + sp_rank."""
    cfg_branches: int
    sp_size: int
    branch: int
    sp_rank: int
    sp_group: object = None     # ranks of my branch
    cfg_group: object = None    # the ranks with my sp_rank, one per branch


def make_layout(world: int, rank: int, cfg_branches: int, new_group: Callable = None) -> ParallelLayout:
    if cfg_branches not in (1, 2) or world % cfg_branches:
        raise ValueError("cfg_branches must be 1 or 2 and divide the world size")
    sp = world // cfg_branches
    lay = ParallelLayout(cfg_branches, sp, rank // sp, rank % sp)
    if new_group is None:
        new_group = dist.new_group
    if world > 1:
        # every rank must create every group, in the same order
        for b in range(cfg_branches):
            g = new_group(list(range(b * sp, (b + 1) * sp)))
            if b == lay.branch:
                lay.sp_group = g
        for s in range(sp):
            g = new_group([b * sp + s for b in range(cfg_branches)])
            if s == lay.sp_rank:
                lay.cfg_group = g
    return lay


def exchange_velocities(v_local: torch.Tensor, layout: ParallelLayout):
    """all-gather of the per-branch velocity over the CFG pair -> (v_cond, v_uncond), identical on both ranks."""
    if layout.cfg_branches == 1:
        return v_local, None
    B = v_local.shape[0]
    buf = torch.empty((2 * B,) + tuple(v_local.shape[1:]), dtype=v_local.dtype, device=v_local.device)
    dist.all_gather_into_tensor(buf, v_local.contiguous(), group=layout.cfg_group)   # concatenated along dim 0
    return buf[:B], buf[B:]


@torch.no_grad()
def run_denoise_parallel(transformer, layout: ParallelLayout, latents, image_latents, prompt_embeds, negative_prompt_embeds,
                         img_shapes, num_inference_steps: int, true_cfg_scale: float = 4.0,
                         sigmas: Optional[Sequence[float]] = None, step_fn: Callable = None, timestep_fn: Callable = None,
                         collect: Optional[list] = None):
    """The denoise loop of pipeline.run_denoise with the two CFG branches on different ranks (and, if the transformer is a
    UlyssesTransformer, each forward spread over the branch's ranks).  `step_fn` defaults to the fused CUDA CFG+Euler kernel."""
    from .pipeline import cfg_euler_step, flowmatch_sigmas, model_timestep
    step_fn = step_fn or cfg_euler_step
    timestep_fn = timestep_fn or model_timestep
    latents = latents.clone()
    B, n, _ = latents.shape
    sig = np.asarray(sigmas, dtype=np.float32) if sigmas is not None else flowmatch_sigmas(num_inference_steps, n)
    embeds = prompt_embeds if layout.branch == 0 else negative_prompt_embeds
    for i in range(num_inference_steps):
        x = torch.cat([latents, image_latents], dim=1)
        ts = timestep_fn(float(sig[i]), B, latents.device)
        v = transformer(hidden_states=x, timestep=ts, encoder_hidden_states=embeds, img_shapes=img_shapes,
                        txt_seq_lens=[embeds.shape[1]] * B, return_dict=False)[0][:, :n].contiguous()
        vc, vu = exchange_velocities(v, layout)
        if collect is not None:
            collect.append((vc.clone(), None if vu is None else vu.clone()))
        latents = step_fn(latents, vc, vu, true_cfg_scale, float(sig[i]), float(sig[i + 1]))
    return latents
