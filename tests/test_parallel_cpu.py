"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard planning, the Ulysses head<->token all-to-all
data movement (checked against single-process attention), and the CFG-pair velocity exchange inside the denoise loop
(checked against the oracle's single-process loop)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

import qie_b200
from oracle import qwen_mmdit_ref as R

PORT = 29641


def _spawn(fn, world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r
    return dict((r[0], r[2]) for r in res)


def _entry(fn, rank, world, q, *args):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(PORT)
    try:
        torch.set_num_threads(2)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = fn(rank, world, *args)
        dist.destroy_process_group()
        q.put((rank, "ok", out))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), None))


# ------------------------------------------------------------------ shard planning (single process)
@pytest.mark.parametrize("img,txt,P", [(8192, 256, 4), (8192, 219, 2), (3072, 427, 8), (8192, 256, 8), (250, 19, 2)])
def test_shard_plan_covers_sequence(img, txt, P):
    plans = [qie_b200.make_shard_plan(img, txt, P, r) for r in range(P)]
    assert sum(p.img_rows for p in plans) == img and sum(p.txt_rows for p in plans) == txt
    for a, b in zip(plans, plans[1:]):
        assert a.img_offset + a.img_rows == b.img_offset and a.txt_offset + a.txt_rows == b.txt_offset
        assert (a.img_pad, a.txt_pad, a.tile_valid) == (b.img_pad, b.txt_pad, b.tile_valid)
    p0 = plans[0]
    assert len(p0.tile_valid) == p0.gathered_rows // 128 and min(p0.tile_valid) >= 1
    assert sum(p0.tile_valid) == img + txt
    # the gathered sequence of the fused exchange is never longer than the single-GPU one (text is not padded per rank)
    assert p0.gathered_rows <= P * p0.img_pad + (txt + 127) // 128 * 128
    rm = p0.rank_major_tiles()
    assert len(rm) == P * p0.rows_pad // 128 and sum(rm) == img + txt


@pytest.mark.parametrize("img,txt,P", [(8192, 256, 4), (8192, 219, 2), (3072, 427, 8), (8192, 256, 8), (250, 19, 2), (700, 150, 3)])
def test_library_shard_plan_equals_python_plan(img, txt, P):
    """qie_sp_shard / qie_sp_tile_valid_host (host-only entry points of libqie.so; what qie_set_peers derives the tile list of
    the gathered sequence from) agree with parallel.make_shard_plan for every rank."""
    import ctypes as C
    from qie_b200 import _lib as L
    lib = qie_b200.lib()
    for r in range(P):
        p = qie_b200.make_shard_plan(img, txt, P, r)
        s, sp = L.Seq(), L.Sp()
        assert lib.qie_sp_shard(3, img, txt, P, r, C.byref(s), C.byref(sp)) == 0
        assert (s.batch, s.img_rows, s.txt_rows, s.img_pad, s.txt_pad) == (3, p.img_rows, p.txt_rows, p.img_pad, p.txt_pad)
        assert (sp.rank, sp.size, sp.img_total, sp.txt_total, sp.img_offset, sp.txt_offset) == (r, P, img, txt, p.img_offset, p.txt_offset)
    buf = (C.c_int * 1024)()
    n = lib.qie_sp_tile_valid_host(img, txt, P, buf, 1024)
    assert tuple(buf[:n]) == p.tile_valid
    assert lib.qie_sp_tile_valid_host(img, txt, P, buf, 2) == -4          # QIE_ENOMEM: no room
    s, sp = L.Seq(), L.Sp()
    assert lib.qie_sp_shard(1, 8192, 257, 2, 0, C.byref(s), C.byref(sp)) == -2     # QIE_ESHAPE, like the Python plan


def test_shard_plan_rejects_bad_splits():
    with pytest.raises(ValueError):
        qie_b200.make_shard_plan(8192, 3, 4, 0)          # a rank without text
    with pytest.raises(ValueError):
        qie_b200.make_shard_plan(8192, 257, 2, 0)        # 129 / 128 text rows straddle a tile boundary


def test_pack_unpack_heads_roundtrip():
    x = torch.arange(6 * 3 * 4 * 128, dtype=torch.float32).view(6, 3 * 4 * 128)
    p = qie_b200.pack_heads(x, 2, 4)
    assert p.shape == (2, 6, 3 * 2 * 128)
    # slab g holds heads [2g, 2g+2) of q, k and v
    assert torch.equal(p[1].view(6, 3, 2, 128), x.view(6, 3, 4, 128)[:, :, 2:4])
    o = torch.arange(2 * 6 * 2 * 128, dtype=torch.float32).view(2, 6, 256)
    u = qie_b200.unpack_heads(o, 2)
    assert torch.equal(u[:, :256], o[0]) and torch.equal(u[:, 256:], o[1])


# ------------------------------------------------------------------ Ulysses data movement over gloo
def _ulysses_worker(rank, world, img, txt, H):
    g = torch.Generator().manual_seed(0)
    qkv_img = torch.randn(img, 3, H, 128, generator=g)
    qkv_txt = torch.randn(txt, 3, H, 128, generator=g)
    plan = qie_b200.make_shard_plan(img, txt, world, rank)
    local = torch.zeros(plan.rows_pad, 3, H, 128)
    local[:plan.img_rows] = qkv_img[plan.img_offset: plan.img_offset + plan.img_rows]
    local[plan.img_pad: plan.img_pad + plan.txt_rows] = qkv_txt[plan.txt_offset: plan.txt_offset + plan.txt_rows]
    send = qie_b200.pack_heads(local.view(plan.rows_pad, -1), world, H)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    hl = H // world
    full = recv.view(world * plan.rows_pad, 3, hl, 128)
    valid = torch.cat([torch.arange(128) < n for n in plan.rank_major_tiles()])    # tile_valid semantics of qie_attn_fwd_tiles
    q, k, v = (full[:, i].transpose(0, 1) for i in range(3))                        # [hl, rows, 128]
    mask = valid[None, None, :].expand(1, full.shape[0], -1)
    o = F.scaled_dot_product_attention(q[None], k[None], v[None], attn_mask=mask)[0].transpose(0, 1)   # [rows, hl, 128]
    o_full = o.reshape(world, plan.rows_pad, hl * 128).contiguous()
    o_recv = torch.empty_like(o_full)
    dist.all_to_all_single(o_recv, o_full)
    mine = qie_b200.unpack_heads(o_recv, world)                                    # [rows_pad, H*128]
    # single-process reference on the whole sequence
    allq = torch.cat([qkv_img, qkv_txt], 0)
    ref = F.scaled_dot_product_attention(*(allq[:, i].transpose(0, 1)[None] for i in range(3)))[0].transpose(0, 1).reshape(img + txt, -1)
    ok_img = torch.allclose(mine[:plan.img_rows], ref[plan.img_offset: plan.img_offset + plan.img_rows], atol=1e-5)
    ok_txt = torch.allclose(mine[plan.img_pad: plan.img_pad + plan.txt_rows],
                            ref[img + plan.txt_offset: img + plan.txt_offset + plan.txt_rows], atol=1e-5)
    return bool(ok_img and ok_txt)


def test_ulysses_all_to_all_matches_single_process_attention():
    res = _spawn(_ulysses_worker, 2, 250, 19, 4)
    assert res == {0: True, 1: True}


# ------------------------------------------------------------------ CFG pair over gloo
def _cfg_worker(rank, world):
    m = R.init_weights_(R.QwenImageTransformer2DModelRef(R.TINY_CONFIG), seed=0).eval()
    shapes = [[(1, 8, 8), (1, 8, 8)]]
    g = torch.Generator().manual_seed(3)
    lat, img_lat = torch.randn(1, 64, 64, generator=g), torch.randn(1, 64, 64, generator=g)
    cond, unc = torch.randn(1, 9, 64, generator=g), torch.randn(1, 5, 64, generator=g)
    layout = qie_b200.make_layout(world, rank, 2)

    def step(latents, vc, vu, scale, s, s_next):                  # oracle stand-ins for the CUDA kernel (test only)
        return R.ref_euler_step(latents, R.ref_cfg_combine(vc, vu, scale), s, s_next)

    def tsf(sigma, B, device):
        return R.ref_timestep_for_model(sigma, torch.float32).expand(B)

    with torch.no_grad():
        got = qie_b200.run_denoise_parallel(m, layout, lat, img_lat, cond, unc, shapes, 3, 4.0,
                                            sigmas=R.ref_flowmatch_sigmas(3, 64), step_fn=step, timestep_fn=tsf)
        ref = R.ref_run_denoise(m, lat, img_lat, cond, shapes, 3, unc, 4.0)
    return (layout.branch, float((got - ref).abs().max()))


def test_cfg_pair_matches_single_process_loop():
    res = _spawn(_cfg_worker, 2)
    assert res[0][0] == 0 and res[1][0] == 1
    assert res[0][1] < 1e-5 and res[1][1] < 1e-5


def test_layout_ranks():
    lay = [qie_b200.make_layout(8, r, 2, new_group=lambda ranks: tuple(ranks)) for r in range(8)]
    assert [l.branch for l in lay] == [0, 0, 0, 0, 1, 1, 1, 1] and [l.sp_rank for l in lay] == [0, 1, 2, 3] * 2
    assert lay[5].sp_group == (4, 5, 6, 7) and lay[5].cfg_group == (1, 5)


# ------------------------------------------------------------------ fused peer-memory exchange: address arithmetic on the host
@pytest.mark.parametrize("P,H,img,txt", [(2, 4, 250, 19), (4, 4, 1000, 300), (8, 8, 1024, 256), (2, 6, 384, 130), (3, 3, 700, 150)])
def test_fused_scatter_addressing_reproduces_single_process_attention(P, H, img, txt):
    """The destination layout the fused epilogues write (scatter_*_reference = the address arithmetic of csrc/gemm.cu and
    csrc/attn.cu with qie_peers): every rank's gathered buffer holds [rank 0 image shard | ... | all text tokens] of its head
    group, attention over it with the plan's tile list equals single-process attention, and the output scatter hands every rank
    the rows of its own tokens — all ranks simulated in one process."""
    g = torch.Generator().manual_seed(P * 100 + H)
    hl = H // P
    full = torch.randn(img + txt, 3, H, 128, generator=g)                      # [img tokens; txt tokens]
    plans = [qie_b200.make_shard_plan(img, txt, P, r) for r in range(P)]
    p0 = plans[0]
    gathered = [torch.zeros(p0.gathered_rows, 3 * hl * 128) for _ in range(P)]
    for p in plans:
        local = torch.zeros(p.rows_pad, 3, H, 128)
        local[:p.img_rows] = full[p.img_offset: p.img_offset + p.img_rows]
        local[p.img_pad: p.img_pad + p.txt_rows] = full[img + p.txt_offset: img + p.txt_offset + p.txt_rows]
        qie_b200.scatter_qkv_reference(local.view(p.rows_pad, -1), gathered, p, H)
    valid = torch.cat([torch.arange(128) < n for n in p0.tile_valid])
    assert int(valid.sum()) == img + txt and valid.numel() == p0.gathered_rows
    ref = F.scaled_dot_product_attention(*(full[:, i].transpose(0, 1)[None] for i in range(3)))[0].transpose(0, 1)   # [S, H, 128]
    attn = [torch.zeros(p0.rows_pad, H * 128) for _ in range(P)]
    for gk in range(P):
        x = gathered[gk].view(p0.gathered_rows, 3, hl, 128)
        # the gathered buffer holds exactly the tokens of the whole sequence, in order, for head group gk
        assert torch.equal(x[valid], full[:, :, gk * hl:(gk + 1) * hl])
        q, k, v = (x[:, i].transpose(0, 1)[None] for i in range(3))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=valid[None, None, None, :])[0].transpose(0, 1)
        qie_b200.scatter_attn_reference(o.reshape(p0.gathered_rows, hl * 128), attn, plans, gk, H)
    for p in plans:
        mine = attn[p.rank].view(p0.rows_pad, H, 128)
        assert torch.allclose(mine[:p.img_rows], ref[p.img_offset: p.img_offset + p.img_rows], atol=1e-5)
        assert torch.allclose(mine[p.img_pad: p.img_pad + p.txt_rows], ref[img + p.txt_offset: img + p.txt_offset + p.txt_rows], atol=1e-5)
        assert not mine[p.img_rows: p.img_pad].any() and not mine[p.img_pad + p.txt_rows:].any()     # pad rows untouched


def test_exchange_bytes_accounting():
    p = qie_b200.make_shard_plan(8192, 256, 4, 1)
    rows, D = p.img_rows + p.txt_rows, 3072
    per_block = rows * 3 * D * 2 * 3 // 4 + (8448 - rows) * (D // 4) * 2
    assert qie_b200.exchange_bytes_per_forward(p, 1, 24, 60) == 60 * per_block + p.img_rows * 64 * 2 * 3 + 60 * 12 * D * 4 // 4 * 3


def test_pure_sequence_parallel_layout_runs_both_cfg_forwards():
    """run_denoise_parallel with one CFG branch (all ranks in ONE sequence-parallel group) must still apply true CFG: both
    forwards run on the group (they were silently dropped)."""
    m = R.init_weights_(R.QwenImageTransformer2DModelRef(R.TINY_CONFIG), seed=0).eval()
    shapes = [[(1, 8, 8), (1, 8, 8)]]
    g = torch.Generator().manual_seed(3)
    lat, img_lat = torch.randn(1, 64, 64, generator=g), torch.randn(1, 64, 64, generator=g)
    cond, unc = torch.randn(1, 9, 64, generator=g), torch.randn(1, 5, 64, generator=g)
    layout = qie_b200.make_layout(1, 0, 1)

    def step(latents, vc, vu, scale, s, s_next):
        return R.ref_euler_step(latents, R.ref_cfg_combine(vc, vu, scale) if vu is not None else vc, s, s_next)

    def tsf(sigma, B, device):
        return R.ref_timestep_for_model(sigma, torch.float32).expand(B)

    with torch.no_grad():
        got = qie_b200.run_denoise_parallel(m, layout, lat, img_lat, cond, unc, shapes, 3, 4.0,
                                            sigmas=R.ref_flowmatch_sigmas(3, 64), step_fn=step, timestep_fn=tsf)
        ref = R.ref_run_denoise(m, lat, img_lat, cond, shapes, 3, unc, 4.0)
        cond_only = R.ref_run_denoise(m, lat, img_lat, cond, shapes, 3, None, 1.0)
    assert float((got - ref).abs().max()) < 1e-5
    assert float((got - cond_only).abs().max()) > 1e-3       # i.e. the uncond branch really contributed


def test_shard_plan_properties_hypothesis():
    """Random (image tokens, text tokens, ranks): a plan either rejects the split (a rank would own an all-padding tile) or
    covers every token exactly once with identical padding on every rank."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 20000), st.integers(1, 1200), st.sampled_from([1, 2, 3, 4, 6, 8]))
    def prop(img, txt, P):
        try:
            plans = [qie_b200.make_shard_plan(img, txt, P, r) for r in range(P)]
        except ValueError:
            return
        assert sum(p.img_rows for p in plans) == img and sum(p.txt_rows for p in plans) == txt
        assert len({(p.img_pad, p.txt_pad, p.tile_valid, p.gathered_rows) for p in plans}) == 1
        p0 = plans[0]
        assert p0.img_pad % 128 == 0 and p0.txt_pad % 128 == 0
        assert all(1 <= v <= 128 for v in p0.tile_valid) and sum(p0.tile_valid) == img + txt
        off_i = off_t = 0
        for p in plans:
            assert (p.img_offset, p.txt_offset) == (off_i, off_t) and p.img_rows >= 1 and p.txt_rows >= 1
            off_i += p.img_rows
            off_t += p.txt_rows

    prop()
