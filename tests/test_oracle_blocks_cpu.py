"""CPU suite: every building block of the oracle against a SECOND, independent statement of the same formula — plain numpy
float64 loops/expressions written from SURVEY.md Appendix A (not from the oracle's code), plus the one independent
implementation that exists in this image (transformers' Qwen2RMSNorm).  Parity with the real diffusers module stays unpinned
(diffusers is absent); what this pins is that the oracle says what Appendix A says, block by block, including the conventions
that are easy to get wrong: [cos | sin] order of the timestep projection, (shift, scale, gate) chunk order, scale-before-shift in
norm_out, adjacent-pair RoPE, text-first concatenation, centred h/w RoPE indices and the text offset."""
import math

import numpy as np
import pytest
import torch

from oracle import qwen_mmdit_ref as R


def f64(t):
    return t.detach().double().numpy()


def np_layernorm(x, eps=1e-6):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps)


def np_rms(x, w, eps=1e-6):
    return x / np.sqrt((x ** 2).mean(-1, keepdims=True) + eps) * w


def np_silu(x):
    return x / (1.0 + np.exp(-x))


def np_gelu_tanh(x):
    return 0.5 * x * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def np_linear(x, lin):
    return x @ f64(lin.weight).T + f64(lin.bias)


def test_timestep_projection_is_cos_then_sin_of_1000t():
    """A.2: f_j = exp(-ln(10000) j / 128); a = 1000 t f_j; emb = [cos a | sin a]."""
    t = torch.tensor([1.0, 0.76953125, 0.02001953125])
    got = f64(R.get_timestep_embedding(t, 256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000))
    j = np.arange(128, dtype=np.float64)
    a = 1000.0 * t.double().numpy()[:, None] * np.exp(-math.log(10000.0) * j / 128.0)[None, :]
    want = np.concatenate([np.cos(a), np.sin(a)], axis=1)
    assert got.shape == (3, 256)
    assert np.abs(got - want).max() < 2e-4       # fp32 sin/cos of arguments up to 1000 rad


def test_timestep_mlp_is_linear_silu_linear():
    m = R.QwenTimestepProjEmbeddings(48)
    t = torch.tensor([0.5, 0.25])
    got = f64(m(t, torch.zeros(1)))
    proj = f64(R.get_timestep_embedding(t, 256, True, 0, 1000))
    want = np_linear(np_silu(np_linear(proj, m.timestep_embedder.linear_1)), m.timestep_embedder.linear_2)
    assert np.abs(got - want).max() < 1e-5


def test_rmsnorm_formula_and_transformers_implementation():
    torch.manual_seed(0)
    n = R.RMSNorm(128, eps=1e-6)
    with torch.no_grad():
        n.weight.copy_(1 + 0.1 * torch.randn(128))
    x = torch.randn(3, 5, 128) * 4
    got = f64(n(x))
    assert np.abs(got - np_rms(f64(x), f64(n.weight))).max() < 1e-5
    qwen2 = pytest.importorskip("transformers.models.qwen2.modeling_qwen2")     # independent implementation in this image
    other = qwen2.Qwen2RMSNorm(128, eps=1e-6)
    with torch.no_grad():
        other.weight.copy_(n.weight)
    assert torch.allclose(n(x), other(x), atol=1e-6, rtol=1e-6)


def test_rope_is_adjacent_pair_complex_rotation():
    """A.4: out[2i] = a cos - b sin, out[2i+1] = a sin + b cos with (a, b) = (x[2i], x[2i+1]); one table for all heads."""
    torch.manual_seed(1)
    x = torch.randn(2, 7, 3, 128)
    theta = torch.rand(7, 64) * 6.0
    freqs = torch.polar(torch.ones_like(theta), theta)
    got = f64(R.apply_rotary_emb_qwen(x, freqs))
    xs, th = f64(x), theta.double().numpy()
    want = np.empty_like(xs)
    for s in range(7):
        for i in range(64):
            a, b = xs[:, s, :, 2 * i], xs[:, s, :, 2 * i + 1]
            want[:, s, :, 2 * i] = a * math.cos(th[s, i]) - b * math.sin(th[s, i])
            want[:, s, :, 2 * i + 1] = a * math.sin(th[s, i]) + b * math.cos(th[s, i])
    assert np.abs(got - want).max() < 1e-5


def test_rope_table_positions():
    """A.5: image k has frame index k, h/w indices -ceil(n/2) .. floor(n/2)-1, row-major over (f, h, w); text token j sits at
    (m+j, m+j, m+j) with m = max over images of max(h//2, w//2); angle = index / 10000^(2p/dim) per axis (16, 56, 56)."""
    rope = R.QwenEmbedRope(10000, [16, 56, 56], scale_rope=True)
    shapes = [(1, 6, 10), (1, 4, 4), (1, 8, 2)]
    img, txt = rope(shapes, [9])

    def angles(idx, dim):
        return idx * (1.0 / 10000.0 ** (np.arange(0, dim, 2, dtype=np.float64) / dim))

    rows = []
    for k, (f, h, w) in enumerate(shapes):
        for fi in range(f):
            for y in range(h):
                for x in range(w):
                    rows.append(np.concatenate([angles(k + fi, 16), angles(y - (h - h // 2), 56), angles(x - (w - w // 2), 56)]))
    want = np.stack(rows)
    got = np.angle(img.numpy().astype(np.complex128))
    assert img.shape == (6 * 10 + 16 + 16, 64)
    assert np.abs(np.exp(1j * want) - np.exp(1j * got)).max() < 1e-4
    m = max(max(h // 2, w // 2) for _, h, w in shapes)          # = 5
    want_t = np.stack([np.concatenate([angles(m + j, 16), angles(m + j, 56), angles(m + j, 56)]) for j in range(9)])
    assert np.abs(np.exp(1j * want_t) - txt.numpy().astype(np.complex128)).max() < 1e-4


def test_feedforward_is_gelu_tanh_mlp():
    torch.manual_seed(2)
    ff = R.FeedForward(32)
    x = torch.randn(2, 5, 32) * 2
    lin1 = ff.net[0].proj
    lin2 = ff.net[2]
    want = np_linear(np_gelu_tanh(np_linear(f64(x), lin1)), lin2)
    assert lin1.out_features == 128 and lin2.in_features == 128          # mult = 4
    assert np.abs(f64(ff(x)) - want).max() < 1e-5


def _np_attention(att, img, txt, img_fr, txt_fr):
    """A.4 in numpy float64: projections, per-head RMSNorm on q/k, RoPE on q/k, TEXT FIRST, softmax(q k^T / sqrt(d)) v."""
    H, Dh = att.heads, att.dim_head

    def heads(x, lin):
        return np_linear(x, lin).reshape(x.shape[0], x.shape[1], H, Dh)

    def rope(x, fr):
        xc = x[..., 0::2] + 1j * x[..., 1::2]
        y = xc * fr.numpy().astype(np.complex128)[None, :, None, :]
        out = np.empty_like(x)
        out[..., 0::2], out[..., 1::2] = y.real, y.imag
        return out

    iq, ik, iv = heads(img, att.to_q), heads(img, att.to_k), heads(img, att.to_v)
    tq, tk, tv = heads(txt, att.add_q_proj), heads(txt, att.add_k_proj), heads(txt, att.add_v_proj)
    iq, ik = np_rms(iq, f64(att.norm_q.weight)), np_rms(ik, f64(att.norm_k.weight))
    tq, tk = np_rms(tq, f64(att.norm_added_q.weight)), np_rms(tk, f64(att.norm_added_k.weight))
    iq, ik, tq, tk = rope(iq, img_fr), rope(ik, img_fr), rope(tq, txt_fr), rope(tk, txt_fr)
    q, k, v = (np.concatenate([a, b], 1) for a, b in ((tq, iq), (tk, ik), (tv, iv)))
    s = np.einsum("bqhd,bkhd->bhqk", q, k) / math.sqrt(Dh)
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    o = np.einsum("bhqk,bkhd->bqhd", p, v).reshape(q.shape[0], q.shape[1], H * Dh)
    T = txt.shape[1]
    return np_linear(o[:, T:], att.to_out[0]), np_linear(o[:, :T], att.to_add_out)


def _small_block():
    torch.manual_seed(3)
    blk = R.QwenImageTransformerBlock(64, 2, 32).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.normal_(0, 0.08)
        for n in (blk.attn.norm_q, blk.attn.norm_k, blk.attn.norm_added_q, blk.attn.norm_added_k):
            n.weight.copy_(1 + 0.1 * torch.randn(32))
    rope = R.QwenEmbedRope(10000, [8, 12, 12], scale_rope=True)
    fr = rope([(1, 4, 6), (1, 2, 4)], [5])
    h, e, temb = torch.randn(2, 32, 64), torch.randn(2, 5, 64), torch.randn(2, 64)
    return blk, fr, h, e, temb


def test_joint_attention_is_text_first_softmax_attention():
    blk, fr, h, e, _ = _small_block()
    with torch.no_grad():
        gi, gt = blk.attn(h, e, fr)
    wi, wt = _np_attention(blk.attn, f64(h), f64(e), fr[0], fr[1])
    assert np.abs(f64(gi) - wi).max() < 1e-5 and np.abs(f64(gt) - wt).max() < 1e-5


def test_block_chunk_order_shift_scale_gate_twice():
    """A.3: mod = Linear(SiLU(temb)) -> [shift1 | scale1 | gate1 | shift2 | scale2 | gate2]; LayerNorm without affine, eps 1e-6."""
    blk, fr, h, e, temb = _small_block()
    with torch.no_grad():
        ge, gh = blk(h, e, temb, fr)
    D = 64
    hn, en, tn = f64(h), f64(e), f64(temb)
    mods = []
    for mod in (blk.img_mod[1], blk.txt_mod[1]):
        m = np_linear(np_silu(tn), mod)
        mods.append([m[:, i * D:(i + 1) * D][:, None, :] for i in range(6)])
    (ish1, isc1, ig1, ish2, isc2, ig2), (tsh1, tsc1, tg1, tsh2, tsc2, tg2) = mods
    ai, at = _np_attention(blk.attn, np_layernorm(hn) * (1 + isc1) + ish1, np_layernorm(en) * (1 + tsc1) + tsh1, fr[0], fr[1])
    hn = hn + ig1 * ai
    en = en + tg1 * at

    def mlp(x, ff):
        return np_linear(np_gelu_tanh(np_linear(x, ff.net[0].proj)), ff.net[2])

    hn = hn + ig2 * mlp(np_layernorm(hn) * (1 + isc2) + ish2, blk.img_mlp)
    en = en + tg2 * mlp(np_layernorm(en) * (1 + tsc2) + tsh2, blk.txt_mlp)
    assert np.abs(f64(gh) - hn).max() < 2e-5 and np.abs(f64(ge) - en).max() < 2e-5


def test_norm_out_is_scale_first_then_shift():
    torch.manual_seed(4)
    n = R.AdaLayerNormContinuous(48, 48)
    x, c = torch.randn(2, 9, 48), torch.randn(2, 48)
    emb = np_linear(np_silu(f64(c)), n.linear)
    scale, shift = emb[:, :48], emb[:, 48:]
    want = np_layernorm(f64(x)) * (1 + scale)[:, None, :] + shift[:, None, :]
    with torch.no_grad():
        assert np.abs(f64(n(x, c)) - want).max() < 1e-5


def test_cfg_combine_and_euler_step_formulas():
    """A.6: comb = u + s (c - u); out = comb * ||c|| / ||comb|| per token over the channels; x <- x + (sigma' - sigma) v."""
    torch.manual_seed(5)
    c, u, x = torch.randn(1, 6, 64), torch.randn(1, 6, 64), torch.randn(1, 6, 64)
    comb = f64(u) + 4.0 * (f64(c) - f64(u))
    want = comb * np.linalg.norm(f64(c), axis=-1, keepdims=True) / np.linalg.norm(comb, axis=-1, keepdims=True)
    assert np.abs(f64(R.ref_cfg_combine(c, u, 4.0)) - want).max() < 1e-5
    got = R.ref_euler_step(x, c, 0.75, 0.5)
    assert np.abs(f64(got) - (f64(x) + (0.5 - 0.75) * f64(c))).max() < 1e-6


def test_sigma_schedule_closed_form():
    """A.8: sigmas = linspace(1, 1/N, N); mu from the linear fit through (256, 0.5), (8192, 0.9); exponential time shift
    exp(mu) / (exp(mu) + (1/s - 1)); stretched so that the last sigma is shift_terminal = 0.02; 0 appended."""
    for n, seq in ((2, 4096), (4, 4096), (8, 1024)):
        s = np.linspace(1.0, 1.0 / n, n)
        m = (0.9 - 0.5) / (8192 - 256)
        mu = seq * m + (0.5 - m * 256)
        s = math.exp(mu) / (math.exp(mu) + (1.0 / s - 1.0))
        one_minus = 1.0 - s
        s = 1.0 - one_minus / (one_minus[-1] / (1.0 - 0.02))
        want = np.concatenate([s, [0.0]])
        assert np.abs(R.ref_flowmatch_sigmas(n, seq) - want).max() < 2e-6
