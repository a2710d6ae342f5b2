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
# peer-memory buffers of the fused exchange (qie_peers in include/qie.h)
# ------------------------------------------------------------------------------------------------
class PeerRankBuffers:
    """What ONE rank of a sequence-parallel group owns: its workspace (the attention-output buffer lives inside), the
    gathered q|k|v buffer of its head group and its barrier flags, all cudaMalloc'ed by libqie so that they can be
    exported over CUDA IPC and written by the other ranks' kernels through NVLink."""

    def __init__(self, ws_bytes: int, gather_bytes: int):
        lib = L.lib()
        self.ptrs, self.handles = [], []
        for n in (ws_bytes, gather_bytes, 256):
            p, h = C.c_void_p(), C.create_string_buffer(64)
            L.check(lib.qie_peer_alloc(n, C.byref(p), h), "qie_peer_alloc")
            if n == ws_bytes and p.value % 1024:      # qie_forward wants a 1 KB aligned workspace; large cudaMallocs are
                raise L.QieError("cudaMalloc returned a workspace that is not 1 KB aligned")
            self.ptrs.append(p.value)
            self.handles.append(h.raw)
        self.ws, self.gather, self.flags = self.ptrs
        self.ws_bytes = ws_bytes

    def free(self):
        for p in self.ptrs:
            L.lib().qie_peer_free(C.c_void_p(p))
        self.ptrs = []


def scatter_qkv_reference(qkv_local: torch.Tensor, gathered: Sequence[torch.Tensor], rank: int, size: int, heads: int) -> None:
    """Host-side statement of what the QKV-GEMM epilogue does with `qie_peers` installed (csrc/gemm.cu, `scat`): head group g of
    q|k|v of MY rows lands in rank g's gathered buffer [size*rows, 3*(H/size)*128] at rows [rank*rows, (rank+1)*rows).
    Equivalent to pack_heads + all_to_all_single of the NCCL form; used by the CPU tests to pin the address arithmetic."""
    rows = qkv_local.shape[0]
    hl = heads // size
    x = qkv_local.view(rows, 3, size, hl * 128)
    for g in range(size):
        gathered[g].view(size, rows, 3, hl * 128)[rank] = x[:, :, g]


def scatter_attn_reference(o_gathered: torch.Tensor, attn_out: Sequence[torch.Tensor], rank: int, size: int, heads: int) -> None:
    """... and what the attention epilogue does: my head group's output rows of rank s's tokens land in rank s's attention
    buffer [rows, H*128] at head columns [rank*(H/size)*128, (rank+1)*(H/size)*128)."""
    hl = heads // size
    rows = o_gathered.shape[0] // size
    for s_ in range(size):
        attn_out[s_][:, rank * hl * 128:(rank + 1) * hl * 128] = o_gathered[s_ * rows:(s_ + 1) * rows]


def make_peers(rank: int, size: int, rows_pad: int, gathers: Sequence[int], attn_outs: Sequence[int], tile_valid: torch.Tensor):
    pr = L.Peers()
    pr.rank, pr.size, pr.rows_pad = rank, size, rows_pad
    for i in range(size):
        pr.qkv_gather[i] = gathers[i]
        pr.attn_out[i] = attn_outs[i]
    pr.tile_valid = tile_valid.data_ptr()
    return pr


# ------------------------------------------------------------------------------------------------
# Ulysses sequence-parallel transformer
# ------------------------------------------------------------------------------------------------
class UlyssesTransformer:
    """Wraps a B200QwenImageTransformer2DModel; same call surface, the work of ONE forward is spread over `group`.

    fused=False: NCCL all-to-alls between the phases (baseline form).
    fused=True : the QKV-GEMM and attention epilogues store straight into the peers' buffers over NVLink (CUDA IPC mapped
                 memory) and the phases are separated by qie_peer_barrier only — no pack / all-to-all / unpack kernels."""

    def __init__(self, transformer, group=None, fused: bool = False):
        self.t = transformer
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if transformer.cfg.num_attention_heads % self.size:
            raise L.QieError(f"{transformer.cfg.num_attention_heads} heads do not split over {self.size} ranks")
        self.config = transformer.config
        self._tile_cache = {}
        self.fused = fused
        self._peer_state = None      # (key, PeerRankBuffers, opened pointers, Peers struct, flag table)
        self._epoch = 0

    # ---- fused path -------------------------------------------------------------------------------------------------
    def _peer_setup(self, plan: "ShardPlan", seq, tiles: torch.Tensor):
        t, lib, P = self.t, L.lib(), self.size
        hl = t.cfg.num_attention_heads // P
        key = (plan.img_pad, plan.txt_pad, P)
        if self._peer_state is not None and self._peer_state[0] == key:
            return self._peer_state
        if self._peer_state is not None:
            raise L.QieError("the fused Ulysses path keeps one shard geometry per wrapper; build a new UlyssesTransformer")
        ws_bytes = lib.qie_workspace_bytes(t._handle, C.byref(seq))
        mine = PeerRankBuffers(ws_bytes, P * plan.rows_pad * 3 * hl * 128 * 2)
        dev = t.device
        blob = torch.frombuffer(bytearray(b"".join(mine.handles)), dtype=torch.uint8).to(dev)
        allb = torch.empty(P * blob.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, blob, group=self.group)
        allb = allb.cpu().view(P, 3, 64)
        ws_p, gather_p, flag_p, opened = [], [], [], []
        for r in range(P):
            if r == self.rank:
                ptrs = (mine.ws, mine.gather, mine.flags)
            else:
                ptrs = []
                for i in range(3):
                    q = C.c_void_p()
                    L.check(lib.qie_peer_open(bytes(allb[r, i].tolist()), C.byref(q)), "qie_peer_open")
                    ptrs.append(q.value)
                    opened.append(q.value)
            ws_p.append(ptrs[0]); gather_p.append(ptrs[1]); flag_p.append(ptrs[2])
        off = lib.qie_workspace_offset(t._handle, C.byref(seq), 1)
        peers = make_peers(self.rank, P, plan.rows_pad, gather_p, [w + off for w in ws_p], tiles)
        L.check(lib.qie_set_peers(t._handle, C.byref(peers), L.cur_stream()), "qie_set_peers")
        flags = (C.c_void_p * P)(*flag_p)
        self._peer_state = (key, mine, opened, peers, flags)
        return self._peer_state

    def _barrier(self, flags):
        self._epoch += 1
        L.check(L.lib().qie_peer_barrier(flags, self.rank, self.size, self._epoch, L.cur_stream()), "qie_peer_barrier")

    def close(self):
        if self._peer_state is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            L.lib().qie_set_peers(self.t._handle, None, None)
            for q in self._peer_state[2]:
                L.lib().qie_peer_close(C.c_void_p(q))
            dist.barrier(group=self.group)
            self._peer_state[1].free()
            self._peer_state = None

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
        D, H, P = t.cfg.inner_dim, t.cfg.num_attention_heads, self.size
        rows = plan.rows_pad
        key = plan.tile_valid
        if key not in self._tile_cache:
            self._tile_cache[key] = torch.tensor(plan.tile_valid, dtype=torch.int32, device=dev)
        tiles = self._tile_cache[key]
        if self.fused:
            return self._call_fused(hidden_states, encoder_hidden_states, timestep, img_shapes, plan, seq, sp, tiles, return_dict)
        ws = t._workspace(seq)
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        ws_bytes = ws.numel() - (base - ws.data_ptr())

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
            out = self._gather_velocity(out_local, S_i)
        out = out.to(hidden_states.dtype) if hidden_states.dtype != torch.bfloat16 else out
        return (out,) if not return_dict else type("Out", (), {"sample": out})()

    def _gather_velocity(self, out_local, S_i):
        """every rank needs the whole velocity for the (replicated) Euler update"""
        t, P, dev = self.t, self.size, self.t.device
        sizes = split_sizes(S_i, P)
        if len(set(sizes)) == 1:
            full = torch.empty(P * sizes[0], t.cfg.out_dim, dtype=torch.bfloat16, device=dev)
            dist.all_gather_into_tensor(full, out_local[0], group=self.group)
            return full.view(1, S_i, t.cfg.out_dim)
        parts = [torch.empty(n, t.cfg.out_dim, dtype=torch.bfloat16, device=dev) for n in sizes]
        dist.all_gather(parts, out_local[0].contiguous(), group=self.group)
        return torch.cat(parts, 0).unsqueeze(0)

    def _call_fused(self, hidden_states, encoder_hidden_states, timestep, img_shapes, plan, seq, sp, tiles, return_dict):
        t, lib, dev = self.t, L.lib(), self.t.device
        _, mine, _, _, flags = self._peer_setup(plan, seq, tiles)
        S_i = hidden_states.shape[1]
        hs = hidden_states[:, plan.img_offset: plan.img_offset + plan.img_rows].to(torch.bfloat16).contiguous()
        enc = encoder_hidden_states[:, plan.txt_offset: plan.txt_offset + plan.txt_rows].to(dev, torch.bfloat16).contiguous()
        ts = timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(1).contiguous()
        shapes = img_shapes[0] if isinstance(img_shapes[0][0], (list, tuple)) else img_shapes
        flat = [int(v) for fhw in shapes for v in fhw]
        shp = (C.c_int * len(flat))(*flat)
        out_local = torch.empty(1, plan.img_rows, t.cfg.out_dim, dtype=torch.bfloat16, device=dev)

        def phase(mask, layer):
            L.check(lib.qie_forward_phase(t._handle, mask, layer, L.ptr(hs), L.ptr(enc), L.ptr(ts), shp, len(flat) // 3,
                                          C.byref(seq), C.byref(sp), L.ptr(out_local), C.c_void_p(mine.ws), mine.ws_bytes, -1,
                                          L.cur_stream()), "qie_forward_phase")

        with torch.cuda.device(dev):
            phase(1, -1)
            for l in range(t.cfg.num_layers):
                phase(2, l)              # adaLN1 + QKV GEMM: the epilogue stores q|k|v of head group g into rank g's gather buffer
                self._barrier(flags)     # everybody's q|k|v has landed in my gather buffer
                phase(4, l)              # attention of my heads over all tokens: the epilogue stores into the token owners' buffers
                self._barrier(flags)     # everybody's heads have landed in my attention buffer (and my gather buffer is free again)
                phase(8, l)              # out-proj, adaLN2, FF on my tokens
            phase(16, -1)
            out = self._gather_velocity(out_local, S_i)
        out = out.to(hidden_states.dtype) if hidden_states.dtype != torch.bfloat16 else out
        return (out,) if not return_dict else type("Out", (), {"sample": out})()


def emulate_fused_ulysses(transformer, size: int, hidden_states, encoder_hidden_states, timestep, img_shapes):
    """Single-GPU emulation of the fused peer-memory path (tests): the `size` ranks live in ONE process on one device, the
    peer tables point at each emulated rank's local buffers, and the ranks' phases run one after the other on one stream
    (stream order replaces qie_peer_barrier; kernels that wait on one another must not share a GPU).  Everything else —
    shard plan, scatter addressing in the QKV-GEMM and attention epilogues, tile list — is the code the real path runs."""
    t, lib, dev = transformer, L.lib(), transformer.device
    S_i, T = hidden_states.shape[1], encoder_hidden_states.shape[1]
    H = t.cfg.num_attention_heads
    hl = H // size
    plans = [make_shard_plan(S_i, T, size, r) for r in range(size)]
    rows = plans[0].rows_pad
    tiles = torch.tensor(plans[0].tile_valid, dtype=torch.int32, device=dev)
    seqs = [L.Seq(1, p.img_rows, p.txt_rows, p.img_pad, p.txt_pad) for p in plans]
    sps = [L.Sp(p.rank, p.size, p.img_total, p.txt_total, p.img_offset, p.txt_offset) for p in plans]
    ws_bytes = max(lib.qie_workspace_bytes(t._handle, C.byref(s)) for s in seqs)
    bufs = [PeerRankBuffers(ws_bytes, size * rows * 3 * hl * 128 * 2) for _ in range(size)]
    off = lib.qie_workspace_offset(t._handle, C.byref(seqs[0]), 1)
    peers = [make_peers(r, size, rows, [b.gather for b in bufs], [b.ws + off for b in bufs], tiles) for r in range(size)]
    ts = timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(1).contiguous()
    shapes = img_shapes[0] if isinstance(img_shapes[0][0], (list, tuple)) else img_shapes
    flat = [int(v) for fhw in shapes for v in fhw]
    shp = (C.c_int * len(flat))(*flat)
    hs = [hidden_states[:, p.img_offset: p.img_offset + p.img_rows].to(torch.bfloat16).contiguous() for p in plans]
    enc = [encoder_hidden_states[:, p.txt_offset: p.txt_offset + p.txt_rows].to(dev, torch.bfloat16).contiguous() for p in plans]
    outs = [torch.empty(1, p.img_rows, t.cfg.out_dim, dtype=torch.bfloat16, device=dev) for p in plans]

    def phase(r, mask, layer):
        L.check(lib.qie_set_peers(t._handle, C.byref(peers[r]), L.cur_stream()), "qie_set_peers")
        L.check(lib.qie_forward_phase(t._handle, mask, layer, L.ptr(hs[r]), L.ptr(enc[r]), L.ptr(ts), shp, len(flat) // 3,
                                      C.byref(seqs[r]), C.byref(sps[r]), L.ptr(outs[r]), C.c_void_p(bufs[r].ws), ws_bytes, -1,
                                      L.cur_stream()), "qie_forward_phase")

    try:
        with torch.cuda.device(dev):
            for r in range(size):
                phase(r, 1, -1)
            for l in range(t.cfg.num_layers):
                for m in (2, 4, 8):
                    for r in range(size):
                        phase(r, m, l)
            for r in range(size):
                phase(r, 16, -1)
            torch.cuda.synchronize()
    finally:
        lib.qie_set_peers(t._handle, None, None)
        torch.cuda.synchronize()
        for b in bufs:
            b.free()
    return torch.cat([o[0] for o in outs], 0).unsqueeze(0)


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
