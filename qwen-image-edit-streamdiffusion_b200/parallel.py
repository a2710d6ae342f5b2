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
# host-side planning (mirror of qie_sp_shard / qie_sp_tile_valid_host in csrc/api.cu)
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
    tile_valid: tuple     # valid rows of every 128-row tile of the GATHERED sequence [P image shards | all text tokens]

    @property
    def rows_pad(self) -> int:
        return self.img_pad + self.txt_pad

    @property
    def gathered_rows(self) -> int:
        """rows of one frame in the gathered layout of the fused exchange: every rank's padded image shard, then ALL text
        tokens contiguously — as long as the single-GPU sequence (no per-rank text padding inside it)"""
        return self.size * self.img_pad + _pad128(self.txt_total)

    @property
    def txt_row0(self) -> int:
        """gathered row of this rank's first text token"""
        return self.size * self.img_pad + self.txt_offset

    def rank_major_tiles(self) -> tuple:
        """tile list of the rank-major layout [rank][img_pad | txt_pad] the NCCL all-to-all form produces"""
        im, tx = split_sizes(self.img_total, self.size), split_sizes(self.txt_total, self.size)
        tiles = []
        for r in range(self.size):
            for n, pad in ((im[r], self.img_pad), (tx[r], self.txt_pad)):
                tiles += [min(128, n - t * 128) for t in range(pad // 128)]
        return tuple(tiles)


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
        tiles += [min(128, im[r] - t * 128) for t in range(img_pad // 128)]
    tiles += [min(128, txt_total - t * 128) for t in range(_pad128(txt_total) // 128)]
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


def exchange_bytes_per_forward(plan: ShardPlan, batch: int, heads: int, layers: int, out_dim: int = 64) -> int:
    """bytes one rank stores into OTHER ranks' memory per forward of the fused exchange: per block the q|k|v of its tokens for
    the other ranks' head groups and the attention output of its head group for the other ranks' tokens, plus its velocity rows"""
    P = plan.size
    rows = plan.img_rows + plan.txt_rows
    D = heads * 128
    qkv = rows * 3 * D * 2 * (P - 1) // P
    att = (plan.img_total + plan.txt_total - rows) * (D // P) * 2
    mod = layers * 12 * D * 4 // P * (P - 1)          # its share of the modulation table, to every other rank
    return batch * (layers * (qkv + att) + plan.img_rows * out_dim * 2 * (P - 1) + mod)


# ------------------------------------------------------------------------------------------------
# peer-memory buffers of the fused exchange (qie_peers in include/qie.h)
# ------------------------------------------------------------------------------------------------
class PeerRankBuffers:
    """What ONE rank of a sequence-parallel group owns: its workspace (the attention-output buffer lives inside), the
    gathered q|k|v buffer of its head group, its velocity buffer, its barrier words and its modulation table, all cudaMalloc'ed by libqie so that
    they can be exported over CUDA IPC and written by the other ranks' kernels through NVLink."""

    def __init__(self, ws_bytes: int, gather_bytes: int, vel_bytes: int, mod_bytes: int):
        lib = L.lib()
        self.ptrs, self.handles = [], []
        for n in (ws_bytes, gather_bytes, vel_bytes, 256, mod_bytes):
            p, h = C.c_void_p(), C.create_string_buffer(64)
            L.check(lib.qie_peer_alloc(n, C.byref(p), h), "qie_peer_alloc")
            if n == ws_bytes and p.value % 1024:      # qie_forward wants a 1 KB aligned workspace; large cudaMallocs are
                raise L.QieError("cudaMalloc returned a workspace that is not 1 KB aligned")
            self.ptrs.append(p.value)
            self.handles.append(h.raw)
        self.ws, self.gather, self.vel, self.flags, self.mod = self.ptrs
        self.sizes = (ws_bytes, gather_bytes, vel_bytes, mod_bytes)

    def free(self):
        for p in self.ptrs:
            L.lib().qie_peer_free(C.c_void_p(p))
        self.ptrs = []


def scatter_qkv_reference(qkv_local: torch.Tensor, gathered: Sequence[torch.Tensor], plan: ShardPlan, heads: int) -> None:
    """Host-side statement of what the QKV-GEMM epilogue does with `qie_peers` installed (csrc/gemm.cu, `scat`): head group g of
    q|k|v of MY valid rows lands in rank g's gathered buffer [gathered_rows, 3*(H/size)*128] — image rows at
    rank*img_pad + i, text rows at size*img_pad + txt_offset + i.  Used by the CPU tests to pin the address arithmetic."""
    rows = qkv_local.shape[0]
    hl = heads // plan.size
    x = qkv_local.view(rows, 3, plan.size, hl * 128)
    for g in range(plan.size):
        dst = gathered[g].view(plan.gathered_rows, 3, hl * 128)
        dst[plan.rank * plan.img_pad: plan.rank * plan.img_pad + plan.img_rows] = x[:plan.img_rows, :, g]
        dst[plan.txt_row0: plan.txt_row0 + plan.txt_rows] = x[plan.img_pad: plan.img_pad + plan.txt_rows, :, g]


def scatter_attn_reference(o_gathered: torch.Tensor, attn_out: Sequence[torch.Tensor], plans: Sequence[ShardPlan], rank: int,
                           heads: int) -> None:
    """... and what the attention epilogue does: my head group's output rows of rank s's tokens land in rank s's attention
    buffer [rows_pad, H*128] at head columns [rank*(H/size)*128, (rank+1)*(H/size)*128)."""
    size = plans[0].size
    hl = heads // size
    cols = slice(rank * hl * 128, (rank + 1) * hl * 128)
    for s_, p in enumerate(plans):
        attn_out[s_][:p.img_rows, cols] = o_gathered[s_ * p.img_pad: s_ * p.img_pad + p.img_rows]
        attn_out[s_][p.img_pad: p.img_pad + p.txt_rows, cols] = o_gathered[p.txt_row0: p.txt_row0 + p.txt_rows]


def make_peers(plan: ShardPlan, batch: int, gathers: Sequence[int], attn_outs: Sequence[int], vels: Sequence[int],
               flags: Sequence[int], mods: Sequence[int]):
    pr = L.Peers()
    pr.rank, pr.size, pr.batch = plan.rank, plan.size, batch
    pr.img_pad, pr.txt_pad, pr.img_total, pr.txt_total = plan.img_pad, plan.txt_pad, plan.img_total, plan.txt_total
    for i in range(plan.size):
        pr.qkv_gather[i] = gathers[i]
        pr.attn_out[i] = attn_outs[i]
        pr.vel[i] = vels[i]
        pr.flags[i] = flags[i]
        pr.mod[i] = mods[i]
    return pr


def _flat_shapes(img_shapes):
    shapes = img_shapes[0] if isinstance(img_shapes[0][0], (list, tuple)) else img_shapes
    return [int(v) for fhw in shapes for v in fhw]


# ------------------------------------------------------------------------------------------------
# Ulysses sequence-parallel transformer
# ------------------------------------------------------------------------------------------------
class _Geometry:
    """per (batch, image tokens, text tokens, img_shapes) state of the fused path: shard plan, C structs, static input /
    output buffers (fixed addresses: the forward is replayed from a CUDA graph) and the captured graph"""

    def __init__(self, t, B, S_i, T, flat, size, rank):
        dev = t.device
        self.B, self.S_i, self.T, self.flat = B, S_i, T, flat
        self.plan = p = make_shard_plan(S_i, T, size, rank)
        self.seq = L.Seq(B, p.img_rows, p.txt_rows, p.img_pad, p.txt_pad)
        self.sp = L.Sp(p.rank, p.size, p.img_total, p.txt_total, p.img_offset, p.txt_offset)
        self.shp = (C.c_int * len(flat))(*flat)
        hl = t.cfg.num_attention_heads // size
        self.ws_bytes = L.lib().qie_workspace_bytes(t._handle, C.byref(self.seq))
        self.gather_bytes = B * p.gathered_rows * 3 * hl * 128 * 2
        self.vel_bytes = B * S_i * t.cfg.out_dim * 2
        self.mod_bytes = B * t.cfg.num_layers * 12 * t.cfg.inner_dim * 4
        self.hs = torch.empty(B, p.img_rows, t.cfg.in_channels, dtype=torch.bfloat16, device=dev)
        self.enc = torch.empty(B, p.txt_rows, t.cfg.joint_attention_dim, dtype=torch.bfloat16, device=dev)
        self.ts = torch.empty(B, dtype=torch.float32, device=dev)
        self.out = torch.empty(B, S_i, t.cfg.out_dim, dtype=torch.bfloat16, device=dev)
        self.graph = None
        self.calls = 0


class UlyssesTransformer:
    """Wraps a B200QwenImageTransformer2DModel; same call surface, the work of ONE forward is spread over `group`.

    fused=False: NCCL all-to-alls between the phases (baseline form, batch 1).
    fused=True : the QKV-GEMM and attention epilogues store straight into the peers' buffers over NVLink (CUDA IPC mapped
                 memory), the END phase stores the velocity rows into every rank's buffer, and the phases are separated by
                 qie_peer_barrier only — no pack / all-to-all / unpack kernels, no NCCL on the data path.  The whole forward is
                 ONE C call (qie_forward_sp), captured in a CUDA graph on its second use (`graphs=True`) and replayed from then
                 on; `pdl=True` launches the per-block kernels with programmatic dependent launch (their prologues overlap the
                 previous kernel's tail).  Any batch size: the frames of a batch stay whole inside the group."""

    def __init__(self, transformer, group=None, fused: bool = False, graphs: bool = True, pdl: bool = True):
        self.t = transformer
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if transformer.cfg.num_attention_heads % self.size:
            raise L.QieError(f"{transformer.cfg.num_attention_heads} heads do not split over {self.size} ranks")
        self.config = transformer.config
        self._tile_cache = {}
        self.fused = fused
        self.graphs = graphs
        if fused and pdl:
            L.check(L.lib().qie_tune(7, 1), "qie_tune")
        self._geo = {}               # geometry key -> _Geometry
        self._bufs = None            # (PeerRankBuffers, opened pointers, per-rank pointer lists)
        self._installed = None       # geometry key whose peers are installed in the handle

    # ---- fused path -------------------------------------------------------------------------------------------------
    def _ensure_buffers(self, geo: _Geometry):
        """(re)allocates the IPC-shared buffers when `geo` needs more room than any geometry before it (collective: every
        rank sees the same geometries in the same order) and exchanges the handles"""
        need = (geo.ws_bytes, geo.gather_bytes, geo.vel_bytes, geo.mod_bytes)
        if self._bufs is not None and all(h >= n for h, n in zip(self._bufs[0].sizes, need)):
            return
        sizes = need if self._bufs is None else tuple(max(h, n) for h, n in zip(self._bufs[0].sizes, need))
        self._release_buffers()
        t, lib, P = self.t, L.lib(), self.size
        mine = PeerRankBuffers(*sizes)
        dev = t.device
        blob = torch.frombuffer(bytearray(b"".join(mine.handles)), dtype=torch.uint8).to(dev)
        allb = torch.empty(P * blob.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, blob, group=self.group)
        allb = allb.cpu().view(P, 5, 64)
        per_rank, opened = [], []
        for r in range(P):
            if r == self.rank:
                ptrs = list(mine.ptrs)
            else:
                ptrs = []
                for i in range(5):
                    q = C.c_void_p()
                    L.check(lib.qie_peer_open(bytes(allb[r, i].tolist()), C.byref(q)), "qie_peer_open")
                    ptrs.append(q.value)
                    opened.append(q.value)
            per_rank.append(ptrs)
        self._bufs = (mine, opened, per_rank)
        for g in self._geo.values():          # captured graphs point into the old buffers
            g.graph = None
            g.calls = 0
        self._installed = None

    def _release_buffers(self):
        if self._bufs is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        L.lib().qie_set_peers(self.t._handle, None, None)
        for q in self._bufs[1]:
            L.lib().qie_peer_close(C.c_void_p(q))
        dist.barrier(group=self.group)
        self._bufs[0].free()
        self._bufs = None
        self._installed = None

    def _install(self, key, geo: _Geometry):
        if self._installed == key:
            return
        t, lib = self.t, L.lib()
        per_rank = self._bufs[2]
        off = lib.qie_workspace_offset(t._handle, C.byref(geo.seq), 1)
        peers = make_peers(geo.plan, geo.B, [p[1] for p in per_rank], [p[0] + off for p in per_rank], [p[2] for p in per_rank],
                           [p[3] for p in per_rank], [p[4] for p in per_rank])
        L.check(lib.qie_set_peers(t._handle, C.byref(peers), L.cur_stream()), "qie_set_peers")
        self._installed = key

    def close(self):
        self._geo = {}
        self._release_buffers()

    def check_barriers(self):
        """raises once a peer barrier of this process has timed out (sticky flag in mapped host memory: no stream is
        synchronised by the read, so it is cheap enough to call after every forward)"""
        if L.lib().qie_peer_barrier_timeouts():
            raise L.QieError("a peer barrier of the sequence-parallel group timed out: a rank is missing or more than 2 s late; "
                             "the results after it are invalid")

    def cache_context(self, name):
        return self.t.cache_context(name)

    def __call__(self, hidden_states, encoder_hidden_states=None, encoder_hidden_states_mask=None, timestep=None,
                 img_shapes=None, txt_seq_lens=None, guidance=None, attention_kwargs=None, return_dict=True, **_):
        if self.fused:
            return self._call_fused(hidden_states, encoder_hidden_states, timestep, img_shapes, return_dict)
        t, lib = self.t, L.lib()
        B, S_i, _ = hidden_states.shape
        if B != 1:
            raise L.QieError("the NCCL form of the sequence-parallel forward is a batch-1 baseline; use fused=True for batches")
        T = encoder_hidden_states.shape[1]
        plan = make_shard_plan(S_i, T, self.size, self.rank)
        dev = t.device
        seq = L.Seq(1, plan.img_rows, plan.txt_rows, plan.img_pad, plan.txt_pad)
        sp = L.Sp(plan.rank, plan.size, plan.img_total, plan.txt_total, plan.img_offset, plan.txt_offset)
        D, H, P = t.cfg.inner_dim, t.cfg.num_attention_heads, self.size
        rows = plan.rows_pad
        key = plan.rank_major_tiles()
        if key not in self._tile_cache:
            self._tile_cache[key] = torch.tensor(key, dtype=torch.int32, device=dev)
        tiles = self._tile_cache[key]
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
        flat = _flat_shapes(img_shapes)
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
                variant = lib.qie_attn_layer_variant(t._handle, l)      # q is pre-scaled in the blocks with bounded scores
                if variant < 0:
                    L.check(variant, "qie_attn_layer_variant")
                L.check(lib.qie_attn_fwd_tiles(L.ptr(q_recv), L.ptr(o_full), P * rows // 128, L.ptr(tiles), hl, variant,
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

    def _call_fused(self, hidden_states, encoder_hidden_states, timestep, img_shapes, return_dict):
        t, lib, dev = self.t, L.lib(), self.t.device
        B, S_i, _ = hidden_states.shape
        T = encoder_hidden_states.shape[1]
        flat = _flat_shapes(img_shapes)
        key = (B, S_i, T, tuple(flat))
        geo = self._geo.get(key)
        if geo is None:
            geo = self._geo[key] = _Geometry(t, B, S_i, T, flat, self.size, self.rank)
        p = geo.plan
        with torch.cuda.device(dev):
            self._ensure_buffers(geo)
            self._install(key, geo)
            mine = self._bufs[0]
            geo.hs.copy_(hidden_states[:, p.img_offset: p.img_offset + p.img_rows])
            geo.enc.copy_(encoder_hidden_states[:, p.txt_offset: p.txt_offset + p.txt_rows])
            geo.ts.copy_(timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(B))

            def run():
                L.check(lib.qie_forward_sp(t._handle, L.ptr(geo.hs), L.ptr(geo.enc), L.ptr(geo.ts), geo.shp, len(flat) // 3,
                                           C.byref(geo.seq), C.byref(geo.sp), L.ptr(geo.out), C.c_void_p(mine.ws), mine.sizes[0],
                                           L.cur_stream()), "qie_forward_sp")

            profiling = getattr(t, "_profiling", False)      # per-kernel CUDA events are recorded by eager launches only
            if geo.graph is not None and not profiling:
                geo.graph.replay()
            elif self.graphs and geo.calls >= 1 and not profiling:
                # second use of this geometry: every lazily built table exists (RoPE rows, tile list, kernel attributes), so the
                # call neither allocates nor synchronises any more and can be recorded
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    run()
                geo.graph = g
                g.replay()
            else:
                run()
            geo.calls += 1
            self.check_barriers()
            out = geo.out.clone()
        out = out.to(hidden_states.dtype) if hidden_states.dtype != torch.bfloat16 else out
        return (out,) if not return_dict else type("Out", (), {"sample": out})()


def emulate_fused_ulysses(transformer, size: int, hidden_states, encoder_hidden_states, timestep, img_shapes):
    """Single-GPU emulation of the fused peer-memory path (tests): the `size` ranks live in ONE process on one device, the
    peer tables point at each emulated rank's local buffers, and the ranks' phases run one after the other on one stream
    (stream order replaces qie_peer_barrier; kernels that wait on one another must not share a GPU).  Everything else —
    shard plan, scatter addressing in the QKV-GEMM / attention / velocity epilogues, tile list — is the code the real path
    runs.  Returns the velocity buffer of every emulated rank ([size][B, S_i, out_dim]; all must be equal)."""
    t, lib, dev = transformer, L.lib(), transformer.device
    B, S_i, T = hidden_states.shape[0], hidden_states.shape[1], encoder_hidden_states.shape[1]
    H = t.cfg.num_attention_heads
    hl = H // size
    plans = [make_shard_plan(S_i, T, size, r) for r in range(size)]
    seqs = [L.Seq(B, p.img_rows, p.txt_rows, p.img_pad, p.txt_pad) for p in plans]
    sps = [L.Sp(p.rank, p.size, p.img_total, p.txt_total, p.img_offset, p.txt_offset) for p in plans]
    ws_bytes = max(lib.qie_workspace_bytes(t._handle, C.byref(s)) for s in seqs)
    vel_bytes = B * S_i * t.cfg.out_dim * 2
    mod_bytes = B * t.cfg.num_layers * 12 * t.cfg.inner_dim * 4
    bufs = [PeerRankBuffers(ws_bytes, B * plans[0].gathered_rows * 3 * hl * 128 * 2, vel_bytes, mod_bytes) for _ in range(size)]
    off = lib.qie_workspace_offset(t._handle, C.byref(seqs[0]), 1)
    peers = [make_peers(plans[r], B, [b.gather for b in bufs], [b.ws + off for b in bufs], [b.vel for b in bufs],
                        [b.flags for b in bufs], [b.mod for b in bufs]) for r in range(size)]
    ts = timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(B).contiguous()
    flat = _flat_shapes(img_shapes)
    shp = (C.c_int * len(flat))(*flat)
    hs = [hidden_states[:, p.img_offset: p.img_offset + p.img_rows].to(torch.bfloat16).contiguous() for p in plans]
    enc = [encoder_hidden_states[:, p.txt_offset: p.txt_offset + p.txt_rows].to(dev, torch.bfloat16).contiguous() for p in plans]

    def phase(r, mask, layer):
        L.check(lib.qie_set_peers(t._handle, C.byref(peers[r]), L.cur_stream()), "qie_set_peers")
        L.check(lib.qie_forward_phase(t._handle, mask, layer, L.ptr(hs[r]), L.ptr(enc[r]), L.ptr(ts), shp, len(flat) // 3,
                                      C.byref(seqs[r]), C.byref(sps[r]), None, C.c_void_p(bufs[r].ws), ws_bytes, -1,
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
                phase(r, 16, -1)           # velocity rows of rank r -> every emulated rank's velocity buffer
            torch.cuda.synchronize()
            outs = []
            for b in bufs:
                o = torch.empty(B, S_i, t.cfg.out_dim, dtype=torch.bfloat16, device=dev)
                _memcpy_d2d(o, b.vel, vel_bytes)
                outs.append(o)
            torch.cuda.synchronize()
    finally:
        lib.qie_set_peers(t._handle, None, None)
        torch.cuda.synchronize()
        for b in bufs:
            b.free()
    return outs


def _memcpy_d2d(dst: torch.Tensor, src_ptr: int, nbytes: int) -> None:
    """device-to-device copy from a raw libqie allocation into a torch tensor, ordered on the current stream"""
    L.check(L.lib().qie_peer_copy(L.ptr(dst), C.c_void_p(src_ptr), nbytes, L.cur_stream()), "qie_peer_copy")


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
    if step_fn is None:          # the fused CUDA CFG+Euler kernel works on bf16 latents (the pipeline's dtype), as run_denoise
        latents, image_latents = latents.to(torch.bfloat16), image_latents.to(torch.bfloat16)
    step_fn = step_fn or cfg_euler_step
    timestep_fn = timestep_fn or model_timestep
    latents = latents.contiguous().clone()
    B, n, _ = latents.shape
    sig = np.asarray(sigmas, dtype=np.float32) if sigmas is not None else flowmatch_sigmas(num_inference_steps, n)
    embeds = prompt_embeds if layout.branch == 0 else negative_prompt_embeds
    do_cfg = true_cfg_scale > 1 and negative_prompt_embeds is not None
    for i in range(num_inference_steps):
        x = torch.cat([latents, image_latents], dim=1)
        ts = timestep_fn(float(sig[i]), B, latents.device)

        def fwd(e):
            return transformer(hidden_states=x, timestep=ts, encoder_hidden_states=e, img_shapes=img_shapes,
                               txt_seq_lens=[e.shape[1]] * B, return_dict=False)[0][:, :n].contiguous()

        v = fwd(embeds)
        if layout.cfg_branches == 1:
            # pure sequence parallelism: both CFG forwards run on this group, one after the other
            vc, vu = v, (fwd(negative_prompt_embeds) if do_cfg else None)
        else:
            vc, vu = exchange_velocities(v, layout)
        if collect is not None:
            collect.append((vc.clone(), None if vu is None else vu.clone()))
        latents = step_fn(latents, vc, vu, true_cfg_scale, float(sig[i]), float(sig[i + 1]))
    return latents
