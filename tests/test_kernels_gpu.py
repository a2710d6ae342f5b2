"""Per-kernel parity (-m gpu): every CUDA kernel, called through the C ABI, against a plain PyTorch fp32
reference of the same op on the same seeded inputs.  Tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

import kernels as K

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def randn(*shape, seed=0, dtype=torch.float32, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, generator=g, device=DEV) * scale).to(dtype)


# ------------------------------------------------------------------ glue kernels
@pytest.mark.parametrize("with_uncond", [False, True])
def test_cfg_euler(with_uncond):
    import qie_b200
    B, n, ch, stride = 2, 1000, 64, 2000      # v holds noise + reference tokens; only the first n are used
    vc = randn(B, stride, ch, seed=1, dtype=torch.bfloat16)
    vu = randn(B, stride, ch, seed=2, dtype=torch.bfloat16) if with_uncond else None
    x = randn(B, n, ch, seed=3, dtype=torch.bfloat16)
    ref = x.float()
    c = vc[:, :n].float()
    if with_uncond:
        u = vu[:, :n].float()
        comb = u + 4.0 * (c - u)
        v = comb * (c.norm(dim=-1, keepdim=True) / comb.norm(dim=-1, keepdim=True))
    else:
        v = c
    ref = (ref + (0.02 - 1.0) * v).to(torch.bfloat16)
    got = qie_b200.cfg_euler_step(x.clone(), vc, vu, 4.0, 1.0, 0.02)
    # one bf16 rounding at the store: <= 1 ulp of bf16 (2^-8 relative)
    assert K.rel_err(got, ref) <= 2 ** -8


@pytest.fixture
def ln_variant(request):
    """qie_tune(3, v): 0 = one warp per row, 1 = warp-per-row streaming ring, 2 = CTA-row form (library default where D is a
    multiple of 1024, else 1).  Restores the library default."""
    lib = K.L.lib()
    prev = lib.qie_tune_get(3)
    K.L.check(lib.qie_tune(3, request.param))
    yield request.param
    K.L.check(lib.qie_tune(3, prev))


@pytest.mark.parametrize("ln_variant", [0, 1, 2], indirect=True)
@pytest.mark.parametrize("D,lens", [(256, [19, 19]), (3072, [19, 19]), (3072, [19, 7]), (1024, [130, 1])])
def test_ln_modulate(D, lens, ln_variant):
    """every adaLN kernel against torch's LayerNorm + modulate; 200 image rows (56 pad rows), text rows that end inside a group of
    four rows of the CTA-row form, one text length per batch element"""
    s = K.L.make_seq_ragged(200, lens) if lens[0] != lens[1] else K.seq(2, 200, lens[0])
    x = randn(K.rows(s), D, seed=4) * 3 + 0.5
    mod = randn(2, 2, 6 * D, seed=5, scale=0.5)
    got = K.ln_modulate(s, x, mod, 2 * 6 * D, 6 * D, 3 * D, 4 * D, D)
    got3 = got.reshape(2, -1, D)
    x3 = x.reshape(2, -1, D)
    for b in range(2):
        for st, (r0, n) in enumerate(((0, s.img_rows), (s.img_pad, lens[b]))):
            ref = F.layer_norm(x3[b, r0:r0 + n], (D,), eps=1e-6) * (1 + mod[b, st, 4 * D:5 * D]) + mod[b, st, 3 * D:4 * D]
            assert K.rel_err(got3[b, r0:r0 + n], ref) <= 2 ** -8
            # pad rows (behind the element's own length) are written as exact zeros
            pad_end = s.img_pad if st == 0 else s.img_pad + s.txt_pad
            assert got3[b, r0 + n:pad_end].abs().max() == 0
    # a second launch on the same stream (row counters re-armed by the first) gives the same bits
    assert torch.equal(K.ln_modulate(s, x, mod, 2 * 6 * D, 6 * D, 3 * D, 4 * D, D), got)


@pytest.mark.parametrize("ln_variant", [1, 2], indirect=True)
@pytest.mark.parametrize("qmode", [1, 2])
def test_ln_modulate_fused_quantiser(qmode, ln_variant):
    """The adaLN kernel's 8-bit shadow output (per-token dynamic quantisation of the bf16-rounded rows, fused into the same pass):
    int8 = exactly torch.round(x / s) with s = amax / 127 (the restated Int8Linear activation quantiser, README.md:136-141; the kernel
    takes the reciprocal product and falls back to the true quotient next to a tie), e4m3 within its 3 mantissa bits."""
    D = 3072
    s = K.seq(1, 300, 77)
    x = randn(K.rows(s), D, seed=14) * 2 + 0.3
    x[:, 11] *= 30                                             # an outlier channel
    mod = randn(1, 2, 6 * D, seed=15, scale=0.5)
    out, out8, sc = K.ln_modulate(s, x, mod, 2 * 6 * D, 6 * D, 0, D, D, fp8=True, qmode=qmode)
    valid = torch.zeros(K.rows(s), dtype=torch.bool, device=DEV)
    valid[:s.img_rows] = True
    valid[s.img_pad:s.img_pad + s.txt_rows] = True
    xf = out.float()[valid]
    qmax = 127.0 if qmode == 2 else 448.0
    amax = xf.abs().amax(dim=1, keepdim=True)
    scale = amax / torch.full_like(amax, qmax)                 # tensor / tensor: a true division (tensor / python float multiplies by 1/x)
    assert torch.equal(sc[valid], scale.squeeze(1))
    if qmode == 2:
        want = torch.round(xf / scale).clamp(-127, 127)
        assert torch.equal(out8[valid].view(torch.int8).float(), want)
    else:
        dq = out8[valid].view(torch.float8_e4m3fn).float() * scale
        assert ((dq - xf).abs().amax(dim=1) <= amax.squeeze(1) * 2 ** -4 + 1e-6).all()
    assert out8[~valid].view(torch.int8).abs().max() == 0 and sc[~valid].abs().max() == 0
    # the W8A8 forward asks for the 8-bit shadow only (out == NULL): same bytes, same scales
    _, only8, only_sc = K.ln_modulate(s, x, mod, 2 * 6 * D, 6 * D, 0, D, D, fp8=True, qmode=qmode, want_bf16=False)
    assert torch.equal(only8, out8) and torch.equal(only_sc, sc)


@pytest.mark.parametrize("B,N,K_,act", [(1, 1000, 256, 0), (2, 6 * 3072, 3072, 1), (8, 514, 512, 1)])
def test_gemv(B, N, K_, act):
    x = randn(B, K_, seed=6)
    w = randn(N, K_, seed=7, dtype=torch.bfloat16, scale=0.05)
    bias = randn(N, seed=8)
    got = K.gemv(x, w, bias, act)
    xin = F.silu(x) if act else x
    ref = xin.double() @ w.double().t() + bias.double()
    assert K.rel_err(got, ref) <= 1e-5      # fp32 accumulation of exact bf16 x fp32 products


def test_timestep_proj():
    t = torch.tensor([1.0, 0.76953125, 0.02001953125], device=DEV)
    got = K.timestep_proj(t)
    half = 128
    f = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=DEV) / half)
    a = 1000 * (t[:, None] * f[None])
    ref = torch.cat([torch.cos(a), torch.sin(a)], dim=-1)
    assert (got - ref).abs().max() <= 2e-4   # sin/cos of arguments up to 1000 in fp32


def test_rmsnorm_pack():
    x = randn(2, 19, 3584, seed=9, dtype=torch.bfloat16, scale=3)
    w = 1 + randn(3584, seed=10, scale=0.02)
    got = K.rmsnorm_pack(x, w, 128)
    xf = x.float()
    ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    assert K.rel_err(got[:, :19], ref) <= 2 ** -8
    assert got[:, 19:].abs().max() == 0


def _rope_ref(x, rope):   # x [rows, H, 128] fp32, rope [rows, 64, 2]
    xr = x.reshape(*x.shape[:-1], 64, 2)
    c, s_ = rope[:, None, :, 0], rope[:, None, :, 1]
    return torch.stack([xr[..., 0] * c - xr[..., 1] * s_, xr[..., 0] * s_ + xr[..., 1] * c], dim=-1).flatten(-2)


def test_qk_norm_rope():
    H = 4
    s = K.seq(1, 130, 19)
    D = H * 128
    qkv = randn(K.rows(s), 3 * D, seed=11, dtype=torch.bfloat16)
    ang = randn(K.rows(s), 64, seed=12, scale=3)
    rope = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
    nw = [1 + randn(128, seed=20 + i, scale=0.1) for i in range(4)]   # img q, img k, txt q, txt k
    ref = qkv.float().clone().reshape(-1, 3, H, 128)
    rp = s.img_pad + s.txt_pad
    for r0, r1, st in ((0, s.img_rows, 0), (s.img_pad, s.img_pad + s.txt_rows, 1)):
        for which in range(2):
            x = ref[r0:r1, which]
            x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * nw[st * 2 + which]
            ref[r0:r1, which] = _rope_ref(x, rope[r0:r1])
    got = K.qk_norm_rope(s, qkv.clone(), rope, nw, H).float().reshape(-1, 3, H, 128)
    valid = torch.zeros(rp, dtype=torch.bool, device=DEV)
    valid[:s.img_rows] = True
    valid[s.img_pad:s.img_pad + s.txt_rows] = True
    assert K.rel_err(got[valid], ref[valid]) <= 2 ** -7
    assert torch.equal(got[:, 2], qkv.float().reshape(-1, 3, H, 128)[:, 2])   # v untouched


# ------------------------------------------------------------------ tcgen05 GEMM
def _gemm_case(s, N, Kd, seed=0):
    a = randn(K.rows(s), Kd, seed=seed, dtype=torch.bfloat16)
    w = [randn(N, Kd, seed=seed + 1 + i, dtype=torch.bfloat16, scale=1 / math.sqrt(Kd)) for i in range(2)]
    b = [randn(N, seed=seed + 3 + i, scale=0.1) for i in range(2)]
    return a, w, b


def _gemm_ref(s, a, w, b):
    ai, at = K.from_joint(s, a.float())
    return ai @ w[0].float().t() + b[0], at @ w[1].float().t() + b[1]


@pytest.mark.parametrize("block_n,cta_group", [(64, 1), (128, 1), (256, 1), (128, 2), (256, 2)])
@pytest.mark.parametrize("B,img,txt", [(1, 256, 128), (2, 200, 19), (1, 384, 300)])
def test_gemm_bf16_epilogue(block_n, cta_group, B, img, txt):
    s = K.seq(B, img, txt)
    N, Kd = 512, 320            # K not a multiple of 64*stages: exercises the ring wrap + TMA K tail
    a, w, b = _gemm_case(s, N, Kd)
    out = torch.full((K.rows(s), N), 7.0, dtype=torch.bfloat16, device=DEV)
    K.gemm(s, a, w, b, out, K.L.EPI_BF16, block_n=block_n, cta_group=cta_group)
    ri, rt = _gemm_ref(s, a, w, b)
    gi, gt = K.from_joint(s, out)
    # fp32 accumulation, one bf16 rounding at the store
    assert K.rel_err(gi, ri) <= 2 ** -7 and K.rel_err(gt, rt) <= 2 ** -7
    pad = out.reshape(B, -1, N)[:, s.img_rows:s.img_pad]
    assert pad.numel() == 0 or pad.abs().max() == 0


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_large_k_many_tiles(cta_group):
    s = K.seq(1, 2048 + 128, 256)
    N, Kd = 3072, 3072
    a, w, b = _gemm_case(s, N, Kd, seed=30)
    out = torch.empty(K.rows(s), N, dtype=torch.bfloat16, device=DEV)
    K.gemm(s, a, w, b, out, K.L.EPI_BF16, cta_group=cta_group)
    ri, rt = _gemm_ref(s, a, w, b)
    gi, gt = K.from_joint(s, out)
    assert K.rel_err(gi, ri) <= 2 ** -7 and K.rel_err(gt, rt) <= 2 ** -7


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_gelu_and_f32(cta_group):
    s = K.seq(1, 384, 100)
    a, w, b = _gemm_case(s, 256, 256, seed=40)
    ri, rt = _gemm_ref(s, a, w, b)
    out = torch.empty(K.rows(s), 256, dtype=torch.bfloat16, device=DEV)
    K.gemm(s, a, w, b, out, K.L.EPI_GELU_BF16, cta_group=cta_group)
    gi, gt = K.from_joint(s, out)
    assert K.rel_err(gi, F.gelu(ri, approximate="tanh")) <= 2 ** -7
    assert K.rel_err(gt, F.gelu(rt, approximate="tanh")) <= 2 ** -7
    out32 = torch.empty(K.rows(s), 256, dtype=torch.float32, device=DEV)
    K.gemm(s, a, w, b, out32, K.L.EPI_F32, cta_group=cta_group)
    gi, gt = K.from_joint(s, out32)
    assert K.rel_err(gi, ri) <= 1e-5 and K.rel_err(gt, rt) <= 1e-5


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_gate_residual(cta_group):
    s = K.seq(2, 256, 60)
    N = 256
    a, w, b = _gemm_case(s, N, 512, seed=50)
    resid0 = randn(K.rows(s), N, seed=51)
    gate = randn(2, 2, 6 * N, seed=52)
    resid = resid0.clone()
    K.gemm(s, a, w, b, resid, K.L.EPI_GATE_RESID_F32, gate=gate[:, :, 2 * N:], gate_bstride=2 * 6 * N,
           gate_sstride=6 * N, cta_group=cta_group)
    ri, rt = _gemm_ref(s, a, w, b)
    r0i, r0t = K.from_joint(s, resid0)
    gi, gt = K.from_joint(s, resid)
    for bb in range(2):
        assert K.rel_err(gi[bb], r0i[bb] + gate[bb, 0, 2 * N:3 * N] * ri[bb]) <= 1e-5
        assert K.rel_err(gt[bb], r0t[bb] + gate[bb, 1, 2 * N:3 * N] * rt[bb]) <= 1e-5
    # pad rows of the residual stream are never touched
    assert torch.equal(resid.reshape(2, -1, N)[:, s.img_pad + s.txt_rows:], resid0.reshape(2, -1, N)[:, s.img_pad + s.txt_rows:])


@pytest.mark.parametrize("epi", ["gate_resid", "gelu", "bf16"])
@pytest.mark.parametrize("img,txt,N,Kd", [(8192, 256, 3072, 1024), (4096 + 200, 219, 3072, 768), (8192, 256, 1792, 3072)])
def test_gemm_split_k_tail(epi, img, txt, N, Kd):
    """Shapes whose persistent schedule ends in a partial wave: the tail tiles are cut into K ranges whose fp32 partials meet
    in scratch and are summed in part order (csrc/gemm.cu, WorkItem).  Checked against the fp32 reference, against the
    un-split schedule of the same kernel, and for run-to-run determinism (no atomics on the data path)."""
    s = K.seq(1, img, txt)
    a, w, b = _gemm_case(s, N, Kd, seed=70)
    rows = K.rows(s)
    gate = randn(1, 2, N, seed=75)
    res0 = randn(rows, N, seed=76)

    def run(split):
        K.L.check(K.L.lib().qie_tune(4, split))
        try:
            if epi == "gate_resid":
                out = res0.clone()
                K.gemm(s, a, w, b, out, K.L.EPI_GATE_RESID_F32, gate=gate, gate_bstride=2 * N, gate_sstride=N, cta_group=2)
            elif epi == "gelu":
                out = torch.empty(rows, N, dtype=torch.bfloat16, device=DEV)
                K.gemm(s, a, w, b, out, K.L.EPI_GELU_BF16, cta_group=2)
            else:
                out = torch.empty(rows, N, dtype=torch.bfloat16, device=DEV)
                K.gemm(s, a, w, b, out, K.L.EPI_BF16, cta_group=2)
            return out
        finally:
            K.L.check(K.L.lib().qie_tune(4, 17))                       # library default: K split of long-K tails, N split elsewhere

    split, again, whole = run(9), run(9), run(0)
    assert torch.equal(split, again)                                   # deterministic
    ri, rt = _gemm_ref(s, a, w, b)
    gi, gt = K.from_joint(s, split.float())
    wi_, wt_ = K.from_joint(s, whole.float())
    if epi == "gate_resid":
        r0i, r0t = K.from_joint(s, res0)
        ri, rt = r0i + gate[:, 0, None, :] * ri, r0t + gate[:, 1, None, :] * rt
    elif epi == "gelu":
        ri, rt = F.gelu(ri, approximate="tanh"), F.gelu(rt, approximate="tanh")
    tol = 1e-5 if epi == "gate_resid" else 2 ** -7
    assert K.rel_err(gi, ri) <= max(tol, 2e-6 * Kd ** 0.5) and K.rel_err(gt, rt) <= max(tol, 2e-6 * Kd ** 0.5)
    # same numbers as the un-split schedule up to the fp32 association of the K ranges
    assert K.rel_err(gi, wi_) <= (5e-6 if epi == "gate_resid" else 2 ** -7) and K.rel_err(gt, wt_) <= (5e-6 if epi == "gate_resid" else 2 ** -7)


@pytest.mark.parametrize("epi", ["gate_resid", "gelu", "bf16", "qkv"])
@pytest.mark.parametrize("img,txt,cta_group", [(8192, 256, 2), (2048, 64, 2), (512, 32, 1), (3000, 219, 2)])
def test_gemm_n_split_tail_is_bit_identical(epi, img, txt, cta_group):
    """N-split tail (csrc/gemm.cu, WorkItem::nhalf): when the partial last wave of the persistent schedule holds at most half as
    many tiles as there are CTAs / CTA pairs, every tail tile is computed as two independent 128-column items (an N = 128 MMA
    into the same accumulator stage, no reduction).  Every output element sees the same K order as in the whole-tile schedule,
    so the result must be BIT-identical to it, for every epilogue incl. the per-head RMSNorm + RoPE one (a half = one head)."""
    s = K.seq(1, img, txt)
    N, Kd = 3072, 512
    H = N // 3 // 128
    a, w, b = _gemm_case(s, N, Kd, seed=77)
    rows = K.rows(s)
    gate = randn(1, 2, N, seed=78)
    res0 = randn(rows, N, seed=79)
    ang = randn(rows, 64, seed=80, scale=3)
    rope = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
    nw = [[1 + randn(128, seed=81 + 2 * st + k, scale=0.1) for k in range(2)] for st in range(2)]
    lib = K.L.lib()
    prev = lib.qie_tune_get(4)

    def run(mode):
        K.L.check(lib.qie_tune(4, mode))
        try:
            if epi == "gate_resid":
                out = res0.clone()
                K.gemm(s, a, w, b, out, K.L.EPI_GATE_RESID_F32, gate=gate, gate_bstride=2 * N, gate_sstride=N, cta_group=cta_group)
            elif epi == "qkv":
                out = torch.empty(rows, N, dtype=torch.bfloat16, device=DEV)
                K.gemm(s, a, w, b, out, K.L.EPI_QKV_NORM_ROPE, rope=rope, qk_norm_w=nw, cta_group=cta_group)
            else:
                out = torch.empty(rows, N, dtype=torch.bfloat16, device=DEV)
                K.gemm(s, a, w, b, out, K.L.EPI_GELU_BF16 if epi == "gelu" else K.L.EPI_BF16, cta_group=cta_group)
            return out
        finally:
            K.L.check(lib.qie_tune(4, prev))

    whole, halves = run(0), run(48)          # 48 = N split of the tail, taking precedence over the K split
    valid = torch.cat([torch.arange(s.img_pad, device=DEV) < s.img_rows, torch.arange(s.txt_pad, device=DEV) < s.txt_rows])
    assert torch.equal(halves[valid], whole[valid])
    if epi == "bf16":                        # and it is the right answer
        ri, rt = _gemm_ref(s, a, w, b)
        gi, gt = K.from_joint(s, halves.float())
        assert K.rel_err(gi, ri) <= 2 ** -7 and K.rel_err(gt, rt) <= 2 ** -7


def test_gemm_compact_single_stream():
    s = K.seq(2, 200, 19)
    N, Kd = 256, 64
    a_img = randn(2 * s.img_pad, Kd, seed=60, dtype=torch.bfloat16)
    w = randn(N, Kd, seed=61, dtype=torch.bfloat16, scale=0.1)
    bias = randn(N, seed=62)
    out = torch.zeros(K.rows(s), N, dtype=torch.float32, device=DEV)
    K.gemm(s, a_img, [w, None], [bias, None], out, K.L.EPI_F32, streams=1, a_compact=1)
    ref = a_img.float().reshape(2, s.img_pad, Kd)[:, :s.img_rows] @ w.float().t() + bias
    gi, gt = K.from_joint(s, out)
    assert K.rel_err(gi, ref) <= 1e-5 and gt.abs().max() == 0
    # text stream only, compact output
    a_j = randn(K.rows(s), Kd, seed=63, dtype=torch.bfloat16)
    outc = torch.empty(2 * s.txt_pad, N, dtype=torch.bfloat16, device=DEV)
    K.gemm(s, a_j, [None, w], [None, bias], outc, K.L.EPI_BF16, streams=2, out_compact=1)
    _, at = K.from_joint(s, a_j.float())
    assert K.rel_err(outc.reshape(2, s.txt_pad, N)[:, :s.txt_rows], at @ w.float().t() + bias) <= 2 ** -7


@pytest.mark.parametrize("block_n,cta_group", [(128, 1), (256, 1), (128, 2), (256, 2)])
def test_gemm_qkv_norm_rope(block_n, cta_group):
    H = 2
    D = H * 128
    s = K.seq(1, 200, 19)
    a, w, b = _gemm_case(s, 3 * D, D, seed=70)
    ang = randn(K.rows(s), 64, seed=71, scale=3)
    rope = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
    nw = [[1 + randn(128, seed=72 + 2 * st + k, scale=0.1) for k in range(2)] for st in range(2)]
    out = torch.empty(K.rows(s), 3 * D, dtype=torch.bfloat16, device=DEV)
    K.gemm(s, a, w, b, out, K.L.EPI_QKV_NORM_ROPE, rope=rope, qk_norm_w=nw, block_n=block_n, cta_group=cta_group)
    refs = _gemm_ref(s, a, w, b)
    ropes = K.from_joint(s, rope.reshape(K.rows(s), 128))
    gots = K.from_joint(s, out)
    for st in range(2):
        r = refs[st][0].reshape(-1, 3, H, 128).clone()
        rp = ropes[st][0].reshape(-1, 64, 2)
        for which in range(2):
            x = r[:, which]
            x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * nw[st][which]
            r[:, which] = _rope_ref(x, rp)
        assert K.rel_err(gots[st][0].reshape(-1, 3, H, 128), r) <= 2 ** -7


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_fp8(cta_group):
    s = K.seq(1, 256, 100)
    N, Kd = 512, 512
    a, w, b = _gemm_case(s, N, Kd, seed=80)
    a8, a_sc = K.quant_rows(a)
    w8, w_sc = [], []
    for wi in w:
        sc = wi.float().abs().amax(dim=1) / 448.0
        w8.append((wi.float() / sc[:, None]).to(torch.float8_e4m3fn).view(torch.uint8))
        w_sc.append(sc.contiguous())
    out = torch.empty(K.rows(s), N, dtype=torch.bfloat16, device=DEV)
    K.gemm(s, a8, w8, b, out, K.L.EPI_BF16, fp8=True, a_scale=a_sc, w_scale=w_sc, cta_group=cta_group)
    # exact reference of the same quantised operands (products of e4m3 values are exact in fp32)
    adq = a8.view(torch.float8_e4m3fn).float() * a_sc[:, None]
    ai, at = K.from_joint(s, adq)
    ri = ai @ (w8[0].view(torch.float8_e4m3fn).float() * w_sc[0][:, None]).t() + b[0]
    rt = at @ (w8[1].view(torch.float8_e4m3fn).float() * w_sc[1][:, None]).t() + b[1]
    gi, gt = K.from_joint(s, out)
    assert K.rel_err(gi, ri) <= 2 ** -7 and K.rel_err(gt, rt) <= 2 ** -7
    # and the activation quantiser itself: |x - dq(q(x))| <= 2^-4 * amax per row (e4m3 has 3 mantissa bits)
    assert ((adq - a.float()).abs().amax(dim=1) <= a.float().abs().amax(dim=1) * 2 ** -4 + 1e-6).all()


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,img,txt,H", [(1, 256, 128, 2), (1, 384, 128, 1), (2, 200, 19, 2), (1, 1024, 219, 3),
                                         (2, 520, 130, 2), (1, 1500, 300, 2)])
@pytest.mark.parametrize("variant", [0, 0x100, 0x20, 0x30, 0x40, 0x108, 0x28])     # CTA-pair kernel (default) / 0x8: single-CTA fallback
def test_attention(B, img, txt, H, variant):
    s = K.seq(B, img, txt)
    D = H * 128
    qkv = randn(K.rows(s), 3 * D, seed=90, dtype=torch.bfloat16)
    got = K.attn(s, qkv, H, variant)
    qi, qt = K.from_joint(s, qkv.float())
    x = torch.cat([qi, qt], dim=1).reshape(B, img + txt, 3, H, 128)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, img + txt, D)
    gi, gt = K.from_joint(s, got)
    g = torch.cat([gi, gt], dim=1)
    # P is rounded to bf16 before P.V and the output is rounded to bf16: 2^-7 of the output scale
    assert K.rel_err(g, ref) <= 2 ** -6, K.rel_err(g, ref)


def test_attention_large_scores_lazy_rescale():
    """Row maxima that keep growing across KV tiles force the lazy O-rescale path."""
    s = K.seq(1, 512, 128)
    H, D = 1, 128
    qkv = randn(K.rows(s), 3 * D, seed=91, dtype=torch.bfloat16)
    ramp = torch.linspace(0.2, 6.0, K.rows(s), device=DEV)[:, None]
    qkv[:, D:2 * D] = (qkv[:, D:2 * D].float() * ramp).to(torch.bfloat16)    # keys grow along the sequence
    qkv[:, :D] = (qkv[:, :D].float() * 3).to(torch.bfloat16)
    got = K.attn(s, qkv, H)
    x = qkv.float().reshape(1, -1, 3, H, 128)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(-1, D)
    assert K.rel_err(got, ref) <= 2 ** -6


@pytest.mark.parametrize("variant", [0, 0x28])
@pytest.mark.parametrize("jump", [3.0, 40.0, 400.0])
def test_attention_score_jumps_between_tiles(variant, jump):
    """Keys whose scale jumps from one KV tile to the next: the lazy rescale path of both kernels (a jump of 400 puts the scores
    of the later tiles hundreds of log2 units above every earlier one)."""
    s = K.seq(1, 768, 128)
    H, D = 1, 128
    qkv = randn(K.rows(s), 3 * D, seed=92, dtype=torch.bfloat16)
    kscale = torch.ones(K.rows(s), 1, device=DEV)
    kscale[512:768] = jump                                   # the third 256-row KV tile carries much larger keys
    qkv[:, D:2 * D] = (qkv[:, D:2 * D].float() * kscale).to(torch.bfloat16)
    qkv[:, :D] = (qkv[:, :D].float() * 4).to(torch.bfloat16)
    got = K.attn(s, qkv, H, variant)
    x = qkv.float().reshape(1, -1, 3, H, 128)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(-1, D)
    assert torch.isfinite(got.float()).all()
    assert K.rel_err(got, ref) <= 2 ** -6


@pytest.mark.parametrize("B,img,txt,H", [(1, 256, 128, 2), (2, 200, 19, 2), (1, 1024, 219, 3), (1, 1500, 300, 2), (1, 384, 128, 1)])
@pytest.mark.parametrize("variant", [0x200, 0x300, 0x220, 0x240])
@pytest.mark.parametrize("wnorm", [1.0, 2.2])
def test_attention_bounded_scores(B, img, txt, H, variant, wnorm):
    """Bounded-score form (0x200): q and k are RMS-normed rows with norm weight `wnorm` (|q.k| * scale * log2e <= 16.3 wnorm^2,
    79 at 2.2 = just under QIE_ATTN_SCORE_BOUND), q arrives pre-multiplied by scale * log2(e); some queries are copies / negated
    copies of keys so that scores sit at both ends of the range.  Same softmax as SDPA on the unscaled q."""
    s = K.seq(B, img, txt)
    D = H * 128
    x = randn(K.rows(s), 3, H, 128, seed=93)
    x[:, :2] = x[:, :2] * torch.rsqrt(x[:, :2].pow(2).mean(-1, keepdim=True)) * wnorm        # |q| = |k| = sqrt(128) * wnorm
    x[5:9, 0] = x[40:44, 1]                  # queries aligned with a key: score = +bound
    x[9:13, 0] = -x[60:64, 1]                # and opposed to one: score = -bound
    c = 0.08838834764831845 * 1.4426950408889634
    xq = x.clone()
    xq[:, 0] *= c
    qkv = xq.reshape(K.rows(s), 3 * D).to(torch.bfloat16)
    got = K.attn(s, qkv, H, variant)
    ref_in = qkv.float().reshape(K.rows(s), 3, H, 128)
    qi, qt = K.from_joint(s, ref_in.reshape(K.rows(s), 3 * D))
    y = torch.cat([qi, qt], dim=1).reshape(B, img + txt, 3, H, 128)
    q, k, v = (y[:, :, i].transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v, scale=0.6931471805599453).transpose(1, 2).reshape(B, img + txt, D)
    gi, gt = K.from_joint(s, got)
    g = torch.cat([gi, gt], dim=1)
    assert torch.isfinite(g.float()).all()
    assert K.rel_err(g, ref) <= 2 ** -6, K.rel_err(g, ref)
    # row-wise too: the rows whose softmax is one spike at the top of the range must not lose to the tolerance of the max norm
    rows = (g.float() - ref).abs().amax(-1) / ref.abs().amax(-1).clamp_min(1e-6)
    assert rows.max() <= 2 ** -5, rows.max()


def test_attention_bounded_equals_online_softmax():
    """Same inputs through the online-softmax kernel (unscaled q) and the bounded-score kernel (pre-scaled q): the two outputs
    agree to the rounding of q * c in bf16."""
    s = K.seq(1, 640, 77)
    H, D = 2, 256
    x = randn(K.rows(s), 3, H, 128, seed=94)
    x[:, :2] = x[:, :2] * torch.rsqrt(x[:, :2].pow(2).mean(-1, keepdim=True))
    a = K.attn(s, x.reshape(-1, 3 * D).to(torch.bfloat16), H, 0)
    xq = x.clone()
    xq[:, 0] *= 0.08838834764831845 * 1.4426950408889634
    b = K.attn(s, xq.reshape(-1, 3 * D).to(torch.bfloat16), H, 0x200)
    assert K.rel_err(b, a.float()) <= 2 ** -6


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_int8_bit_exact_against_int8_oracle(cta_group):
    """INT8 W8A8 (tcgen05 kind::i8, int32 accumulate): the activation quantiser reproduces the restated Int8Linear oracle's
    integers exactly, the integer accumulation is exact, and the dequantised output equals the oracle's up to the final
    bf16 rounding (README.md:136-141 "quantize + matmul + dequantize")."""
    from oracle import qwen_mmdit_ref as R
    s = K.seq(1, 256, 100)
    N, Kd = 512, 512
    a, w, b = _gemm_case(s, N, Kd, seed=85)
    a[:, 7] *= 20                                             # an outlier channel
    a8, a_sc = K.quant_rows(a, qmode=2)
    xf = a.float()
    s_x = xf.abs().amax(dim=1, keepdim=True) / 127.0
    assert torch.equal(a8.view(torch.int8).float(), torch.round(xf / s_x).clamp(-127, 127))    # bit-exact integers
    w8, w_sc = [], []
    for wi in w:      # offline weight quantisation exactly as the oracle does it (on the host: torch-CUDA divides by a
        wc = wi.float().cpu()   # constant via a reciprocal multiply, which can flip a rounding)
        sc = wc.abs().amax(dim=1) / 127.0
        w8.append(torch.round(wc / sc[:, None]).clamp(-127, 127).to(torch.int8).view(torch.uint8).to(DEV))
        w_sc.append(sc.contiguous().to(DEV))
    out = torch.empty(K.rows(s), N, dtype=torch.float32, device=DEV)
    K.gemm(s, a8, w8, b, out, K.L.EPI_F32, fp8=2, a_scale=a_sc, w_scale=w_sc, cta_group=cta_group)
    ai, at = K.from_joint(s, xf)
    gi, gt = K.from_joint(s, out)
    ri = R.ref_int8_linear(ai[0].cpu(), w[0].float().cpu(), b[0].cpu()).to(DEV)
    rt = R.ref_int8_linear(at[0].cpu(), w[1].float().cpu(), b[1].cpu()).to(DEV)
    # same integers, same int32 sums; only the order of the two fp32 scale multiplies may differ by 1 ulp
    assert K.rel_err(gi[0], ri) <= 1e-6 and K.rel_err(gt[0], rt) <= 1e-6
