"""CPU suite (-m "not gpu"): the oracle against its golden vectors / known answers / algebraic identities, the host
logic of the library (scheduler, RoPE table, sequence layout), and the C-ABI surface of libqie.so (load + symbols only;
no compute call needs a GPU here)."""
import ctypes as C
import math
import re
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import qie_b200
from qie_b200 import _lib as L
from oracle import qwen_mmdit_ref as R

ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------------------------ scheduler known answers
@pytest.mark.parametrize("key,n,seq", [("n2_s4096", 2, 4096), ("n4_s4096", 4, 4096), ("n8_s4096", 8, 4096),
                                       ("n4_s1024", 4, 1024), ("n4_s256", 4, 256)])
def test_sigma_tables_oracle_and_library(key, n, seq):
    gold = np.load(GOLD / "sigmas_a8.npz")[key]
    assert np.allclose(R.ref_flowmatch_sigmas(n, seq), gold, atol=2e-6)
    assert np.allclose(qie_b200.flowmatch_sigmas(n, seq), gold, atol=2e-6)          # host entry point of libqie.so
    assert np.allclose(qie_b200.flowmatch_sigmas(n, seq), R.ref_flowmatch_sigmas(n, seq), atol=1e-6)


def test_calculate_shift_known_answers():
    g = np.load(GOLD / "sigmas_a8.npz")
    assert abs(R.calculate_shift(4096) - g["mu_4096"]) < 1e-6
    assert abs(R.calculate_shift(1024) - g["mu_1024"]) < 1e-6
    assert abs(R.calculate_shift(256) - g["mu_256"]) < 1e-6


def test_timestep_rounding_chain():
    # t is cast to bf16 before /1000 -> 766.709 -> 768 -> 0.76953125 ; 20 -> 0.02001953125
    assert R.ref_timestep_for_model(0.766709, torch.bfloat16).float().item() == pytest.approx(0.76953125)
    assert R.ref_timestep_for_model(0.02, torch.bfloat16).float().item() == pytest.approx(0.02001953125)
    assert qie_b200.model_timestep(0.766709, 1, "cpu").float().item() == pytest.approx(0.76953125)


# ------------------------------------------------------------------ oracle regression vectors (configs[0], tiny model)
def _tiny():
    m = R.init_weights_(R.QwenImageTransformer2DModelRef(R.TINY_CONFIG), seed=0).eval()
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    hidden, enc = R.make_inputs(R.TINY_CONFIG, shapes, 19, seed=1)
    return m, shapes, hidden, enc


def test_oracle_matches_golden_velocity_and_denoise():
    g = np.load(GOLD / "tiny_oracle.npz")
    m, shapes, hidden, enc = _tiny()
    with torch.no_grad():
        v = m(hidden, enc, None, torch.tensor([1.0]), shapes, [19])[0]
        lat = R.ref_run_denoise(m, hidden[:, :256], hidden[:, 256:], enc, shapes, 2)
        lat_cfg = R.ref_run_denoise(m, hidden[:, :256], hidden[:, 256:], enc, shapes, 4, enc[:, :11] * 0.5, 4.0)
    assert np.allclose(v.numpy()[0, ::16], g["velocity"], atol=2e-5, rtol=1e-4)
    assert np.allclose(lat.numpy()[0, ::16], g["final_2step"], atol=5e-5, rtol=1e-4)
    assert np.allclose(lat_cfg.numpy()[0, ::16], g["final_4step_cfg"], atol=2e-4, rtol=1e-3)


def test_oracle_is_deterministic_and_fp64_consistent():
    m, shapes, hidden, enc = _tiny()
    with torch.no_grad():
        a = m(hidden, enc, None, torch.tensor([0.5]), shapes, [19])[0]
        b = m(hidden, enc, None, torch.tensor([0.5]), shapes, [19])[0]
        m64 = m.double()
        m64.pos_embed.pos_freqs = m64.pos_embed.pos_freqs.to(torch.complex128)
        m64.pos_embed.neg_freqs = m64.pos_embed.neg_freqs.to(torch.complex128)
        c = m64(hidden.double(), enc.double(), None, torch.tensor([0.5], dtype=torch.float64), shapes, [19])[0]
    assert torch.equal(a, b)
    assert ((a.double() - c).abs().max() / c.abs().max()).item() < 1e-4


# ------------------------------------------------------------------ RoPE invariants + library table
def test_rope_invariants():
    g = np.load(GOLD / "tiny_oracle.npz")
    rope = R.QwenEmbedRope(10000, [16, 56, 56], scale_rope=True)
    img, txt = rope([(1, 64, 64), (1, 64, 64)], [256])
    assert img.shape == (8192, 64) and txt.shape == (256, 64)
    assert torch.allclose(img.abs(), torch.ones(8192, 64), atol=1e-6)             # unit modulus
    # frame index of image k is k: frame axis (first 8 pairs) identical inside an image, differs between images
    assert torch.equal(img[0, :8], img[4095, :8]) and not torch.equal(img[0, :8], img[4096, :8])
    # centred h/w: row 32 / col 32 have index 0 -> freq = 1+0j on those axes
    tok = 32 * 64 + 32
    assert torch.allclose(img[tok, 8:], torch.ones(56, dtype=torch.complex64), atol=1e-6)
    # text offset = max(h//2, w//2) = 32 on all three axes
    assert torch.allclose(txt[0], rope.pos_freqs[32], atol=0)
    tiny = R.QwenEmbedRope(10000, [8, 12, 12], scale_rope=True)
    ti, tt = tiny([(1, 16, 16), (1, 16, 16)], [19])
    assert np.allclose(torch.view_as_real(ti).numpy()[::37], g["rope_img"]) and np.allclose(torch.view_as_real(tt).numpy(), g["rope_txt"])


@pytest.mark.parametrize("shapes,T", [([(1, 64, 64), (1, 64, 64)], 219), ([(1, 10, 12), (1, 16, 8), (1, 7, 9)], 33)])
def test_library_rope_table_matches_oracle(shapes, T):
    """qie_rope_table_host (host C++) == QwenEmbedRope of the oracle, in the joint [img; txt] layout."""
    n_img = sum(f * h * w for f, h, w in shapes)
    seq = qie_b200.make_seq(1, n_img, T)
    cfg = L.ModelCfg(1, 1, 128, 64, 64, 64, (C.c_int * 3)(16, 56, 56))
    flat = [v for s in shapes for v in s]
    buf = (C.c_float * ((seq.img_pad + seq.txt_pad) * 128))()
    L.check(L.lib().qie_rope_table_host(C.byref(cfg), (C.c_int * len(flat))(*flat), len(shapes), C.byref(seq), buf))
    tab = torch.frombuffer(buf, dtype=torch.float32).reshape(-1, 64, 2)
    img, txt = R.QwenEmbedRope(10000, [16, 56, 56], scale_rope=True)(list(shapes), [T])
    assert torch.allclose(tab[:n_img], torch.view_as_real(img), atol=2e-6)
    assert torch.allclose(tab[seq.img_pad:seq.img_pad + T], torch.view_as_real(txt), atol=2e-6)
    pad = tab[n_img:seq.img_pad]
    assert pad.numel() == 0 or (torch.equal(pad[..., 0], torch.ones_like(pad[..., 0])) and pad[..., 1].abs().max() == 0)


# ------------------------------------------------------------------ algebraic identities of the oracle (SURVEY §8c iii, v)
def test_zero_modulation_reduces_to_adaln_projection():
    m, shapes, hidden, enc = _tiny()
    with torch.no_grad():
        for blk in m.transformer_blocks:
            for mod in (blk.img_mod[1], blk.txt_mod[1]):
                mod.weight.zero_(); mod.bias.zero_()
        out = m(hidden, enc, None, torch.tensor([1.0]), shapes, [19])[0]
        h = m.img_in(hidden)
        temb = m.time_text_embed(torch.tensor([1.0]), h)
        ref = m.proj_out(m.norm_out(h, temb))
    assert torch.allclose(out, ref, atol=1e-5)


def test_text_permutation_leaves_image_output_unchanged():
    """Joint attention is permutation-invariant over keys once each token carries its RoPE row — the property the CUDA
    path relies on to lay the sequence out as [img; txt] instead of the reference's cat([txt, img])."""
    m, shapes, hidden, enc = _tiny()
    blk = m.transformer_blocks[0]
    with torch.no_grad():
        h, e = m.img_in(hidden), m.txt_in(m.txt_norm(enc))
        temb = m.time_text_embed(torch.tensor([0.7]), h)
        fr = m.pos_embed(shapes, [19])
        perm = torch.randperm(19, generator=torch.Generator().manual_seed(0))
        e1, h1 = blk(h, e, temb, fr)
        e2, h2 = blk(h, e[:, perm], temb, (fr[0], fr[1][perm]))
    assert torch.allclose(h1, h2, atol=2e-5) and torch.allclose(e1[:, perm], e2, atol=2e-5)


def test_cfg_and_euler_identities():
    g = torch.Generator().manual_seed(0)
    c, u, x = torch.randn(2, 50, 64, generator=g), torch.randn(2, 50, 64, generator=g), torch.randn(2, 50, 64, generator=g)
    assert torch.allclose(R.ref_cfg_combine(c, u, 1.0), c, atol=1e-6)                      # scale 1 -> cond
    v = R.ref_cfg_combine(c, u, 4.0)
    assert torch.allclose(v.norm(dim=-1), c.norm(dim=-1), rtol=1e-5)                      # norm rescale
    assert torch.equal(R.ref_euler_step(x, v, 0.5, 0.5), x)                               # sigma' == sigma -> identity
    assert torch.allclose(R.ref_euler_step(x, v, 1.0, 0.0), x - v)


def test_int8_and_fp8_linear_oracles():
    g = torch.Generator().manual_seed(1)
    x, w, b = torch.randn(64, 256, generator=g), torch.randn(96, 256, generator=g) / 16, torch.randn(96, generator=g)
    exact = x @ w.t() + b
    e8 = (R.ref_int8_linear(x, w, b) - exact).abs().max() / exact.abs().max()
    ef = (R.ref_fp8_linear(x, w, b) - exact).abs().max() / exact.abs().max()
    assert e8 < 2e-2 and ef < 6e-2
    # exactly representable operands -> exact result
    xi, wi = torch.randint(-127, 128, (8, 32)).float(), torch.randint(-127, 128, (16, 32)).float()
    xi[:, 0], wi[:, 0] = 127, 127
    assert torch.equal(R.ref_int8_linear(xi, wi, None), xi @ wi.t())


# ------------------------------------------------------------------ host logic of the library
def test_make_seq_padding_and_errors():
    s = qie_b200.make_seq(2, 8192, 219)
    assert (s.batch, s.img_rows, s.txt_rows, s.img_pad, s.txt_pad) == (2, 8192, 219, 8192, 256)
    s = qie_b200.make_seq(1, 4070, 1)
    assert s.img_pad == 4096 and s.txt_pad == 128
    with pytest.raises(qie_b200.QieError, match="qie_make_seq"):
        qie_b200.make_seq(0, 10, 10)
    with pytest.raises(qie_b200.QieError):
        qie_b200.make_seq(1, 10, 0)


def test_error_convention_null_pointer_and_message():
    lib = L.lib()
    rc = lib.qie_cfg_euler_step(None, None, None, 4.0, 1.0, 0.5, 1, 16, 64, 16, None)
    assert rc == -1 and b"null pointer" in lib.qie_last_error()
    rc = lib.qie_gemm(None, None, None)
    assert rc == -1
    s = qie_b200.make_seq(1, 128, 128)
    assert lib.qie_workspace_bytes(None, C.byref(s)) == 0


def test_abi_exports_every_declared_symbol():
    """Every function include/qie.h declares is exported by libqie.so and bound in _lib.SYMBOLS."""
    header = (ROOT / "include" / "qie.h").read_text()
    declared = set(re.findall(r"\b(qie_[a-z0-9_]+)\s*\(", header))
    declared -= {"qie_status"}
    lib = C.CDLL(str(qie_b200.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in qie.h but not exported"
        assert name in L.SYMBOLS, f"{name} not bound in _lib.SYMBOLS"
    assert lib.qie_version() == 1


def test_struct_layouts_match_header_sizes():
    # ctypes mirrors of the C structs: sizes follow from the field lists in include/qie.h
    assert C.sizeof(L.Seq) == (5 + 8) * 4
    assert C.sizeof(L.ModelCfg) == 9 * 4
    assert C.sizeof(L.BlockWeights) == 18 * 2 * 8
    assert C.sizeof(L.Weights) == 16 * 8


def test_no_cpu_fallback_when_library_missing(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", tmp_path / "absent.so")
    with pytest.raises(qie_b200.QieError, match="no CPU"):
        L.lib()


def test_product_path_does_not_import_the_oracle():
    pkg = ROOT / "qwen-image-edit-streamdiffusion_b200"
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if f.is_file() and f.suffix in (".py", ".cu", ".cuh"):
            assert "oracle" not in f.read_text().replace("int8 oracle", ""), f


def test_merge_lora_host_math_and_naming(tmp_path):
    """N2 host logic: W' = W + scale * alpha/r * B@A for PEFT and kohya key styles; safetensors round trip of the keys."""
    from safetensors.torch import save_file, load_file
    g = torch.Generator().manual_seed(0)
    sd = {"transformer_blocks.0.attn.to_q.weight": torch.randn(16, 8, generator=g),
          "transformer_blocks.0.attn.to_q.bias": torch.randn(16, generator=g),
          "transformer_blocks.0.img_mlp.net.2.weight": torch.randn(8, 32, generator=g)}
    A1, B1 = torch.randn(4, 8, generator=g), torch.randn(16, 4, generator=g)
    A2, B2 = torch.randn(2, 32, generator=g), torch.randn(8, 2, generator=g)
    lora = {"transformer.transformer_blocks.0.attn.to_q.lora_A.weight": A1,
            "transformer.transformer_blocks.0.attn.to_q.lora_B.weight": B1,
            "transformer_blocks.0.img_mlp.net.2.lora_down.weight": A2,
            "transformer_blocks.0.img_mlp.net.2.lora_up.weight": B2,
            "transformer_blocks.0.img_mlp.net.2.alpha": torch.tensor(1.0)}
    save_file(lora, str(tmp_path / "lora.safetensors"))
    m = qie_b200.merge_lora(sd, load_file(str(tmp_path / "lora.safetensors")), scale=0.5)
    assert torch.allclose(m["transformer_blocks.0.attn.to_q.weight"], sd["transformer_blocks.0.attn.to_q.weight"] + 0.5 * B1 @ A1)
    assert torch.allclose(m["transformer_blocks.0.img_mlp.net.2.weight"],
                          sd["transformer_blocks.0.img_mlp.net.2.weight"] + 0.5 * (1.0 / 2) * B2 @ A2)
    assert torch.equal(m["transformer_blocks.0.attn.to_q.bias"], sd["transformer_blocks.0.attn.to_q.bias"])
    with pytest.raises(qie_b200.QieError, match="not a Linear"):
        qie_b200.merge_lora(sd, {"nope.lora_A.weight": A1, "nope.lora_B.weight": B1})


# ------------------------------------------------------------------ next rows N3 / N4
def test_pack_latents_index_formula_and_round_trip():
    """_pack_latents: channel index of a packed token = c*4 + dy*2 + dx; unpack is its inverse; the VAE
    normalisation round-trips."""
    g = torch.Generator().manual_seed(5)
    B, C, h, w = 2, 16, 6, 8
    z = torch.randn(B, C, h, w, generator=g)
    tok = R.ref_pack_latents(z)
    assert tok.shape == (B, (h // 2) * (w // 2), 4 * C)
    for (b, c, y, x) in [(0, 0, 0, 0), (1, 3, 5, 7), (0, 15, 2, 5), (1, 7, 3, 0)]:
        assert tok[b, (y // 2) * (w // 2) + x // 2, c * 4 + (y % 2) * 2 + (x % 2)] == z[b, c, y, x]
    assert torch.equal(R.ref_unpack_latents(tok, h, w)[:, :, 0], z)
    mean, std = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    back = R.ref_unpack_latents(R.ref_pack_latents(z, mean, std), h, w, mean, std)[:, :, 0]
    assert torch.allclose(back, z, atol=1e-5)


def test_stream_prepare_latent_follows_the_reference_sketch():
    """qwen_realtime.py:201-224: key frame every keyframe_interval frames (and when there is no previous latent)."""
    noise, prev = torch.ones(1, 4, 8), torch.full((1, 4, 8), 2.0)
    lat, key = R.ref_stream_prepare_latent(None, noise, 3)
    assert key and torch.equal(lat, noise)
    lat, key = R.ref_stream_prepare_latent(prev, noise, 20)
    assert key and torch.equal(lat, noise)
    lat, key = R.ref_stream_prepare_latent(prev, noise, 21)
    assert not key and torch.allclose(lat, prev + 0.05 * noise)


def test_streaming_denoiser_host_logic_matches_oracle_schedule():
    """StreamingDenoiser (host side of N4) on CPU with an injected denoise function: which frames are key frames, where the
    schedule is entered, and that the previous latent feeds the next frame — against ref_stream_prepare_latent."""
    import qie_b200
    calls = []

    def fake_denoise(t, start, img_lat, cond, shapes, steps, unc, scale, begin_index=0):
        calls.append((start.clone(), begin_index))
        return start * 0.5 + 1.0

    sd = qie_b200.StreamingDenoiser(None, None, None, num_inference_steps=4, keyframe_interval=3, stream_steps=1,
                                    noise_strength=0.05, denoise_fn=fake_denoise)
    g = torch.Generator().manual_seed(0)
    prev = None
    for f in range(7):
        noise = torch.randn(1, 4, 8, generator=g)
        want, key = R.ref_stream_prepare_latent(prev, noise, f, 3, 0.05)
        out = sd.process_frame(None, noise)
        start, begin = calls[-1]
        assert torch.allclose(start, want) and begin == (0 if key else 3) and sd.is_keyframe == key
        assert sd.forwards_per_frame() == (4 if key else 1)
        prev = out
    assert sd.frame_count == 7


# ------------------------------------------------------------------ properties over random shapes (hypothesis)
def test_sigma_schedule_properties_library_vs_oracle():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(2, 50), seq=st.integers(16, 16384))
    def check(n, seq):
        lib_s, ref_s = qie_b200.flowmatch_sigmas(n, seq), R.ref_flowmatch_sigmas(n, seq)
        assert lib_s.shape == (n + 1,) and np.allclose(lib_s, ref_s, atol=2e-6)
        assert lib_s[-1] == 0.0 and abs(lib_s[-2] - 0.02) < 1e-6                    # stretched to shift_terminal, then 0 appended
        assert np.all(np.diff(lib_s) < 0)                                          # strictly decreasing
        assert abs(lib_s[0] - 1.0) < 1e-6

    check()
    # one step: the reference scheduler's stretch is 0/0 (NaN schedule, silently); the host wrapper refuses instead
    with np.errstate(invalid="ignore"):
        assert np.isnan(R.ref_flowmatch_sigmas(1, 4096)[0])
    with pytest.raises(qie_b200.QieError, match="at least 2 steps"):
        qie_b200.flowmatch_sigmas(1, 4096)
    # ... and so does the C entry point itself (a caller that binds libqie directly must not get NaN sigmas with QIE_OK)
    import ctypes as C
    buf = (C.c_float * 2)()
    lib = qie_b200.lib()
    assert lib.qie_flowmatch_sigmas(1, 4096, buf) == -1 and b"at least 2 steps" in lib.qie_last_error()


def test_rope_table_library_vs_oracle_random_image_lists():
    """qie_rope_table_host against QwenEmbedRope for random lists of 1-3 images with odd / even grids and ragged text."""
    from hypothesis import given, settings, strategies as st

    cfg = L.ModelCfg(1, 1, 128, 64, 64, 64, (C.c_int * 3)(16, 56, 56))
    rope = R.QwenEmbedRope(10000, [16, 56, 56], scale_rope=True)
    grid = st.tuples(st.just(1), st.integers(1, 24), st.integers(1, 24))

    @settings(max_examples=25, deadline=None)
    @given(shapes=st.lists(grid, min_size=1, max_size=3), T=st.integers(1, 300))
    def check(shapes, T):
        n_img = sum(f * h * w for f, h, w in shapes)
        seq = qie_b200.make_seq(1, n_img, T)
        flat = [v for s in shapes for v in s]
        buf = (C.c_float * ((seq.img_pad + seq.txt_pad) * 128))()
        L.check(L.lib().qie_rope_table_host(C.byref(cfg), (C.c_int * len(flat))(*flat), len(shapes), C.byref(seq), buf))
        tab = torch.frombuffer(buf, dtype=torch.float32).reshape(-1, 64, 2)
        img, txt = rope([tuple(s) for s in shapes], [T])
        assert torch.allclose(tab[:n_img], torch.view_as_real(img), atol=3e-6)
        assert torch.allclose(tab[seq.img_pad:seq.img_pad + T], torch.view_as_real(txt), atol=3e-6)

    check()


# ------------------------------------------------------------------ the real upstream, whenever it is importable
def _have_diffusers():
    try:
        import diffusers  # noqa: F401
        return True
    except Exception:
        return False


_DIFF_FIXTURE = GOLD / "diffusers_tiny.npz"
_needs_upstream = pytest.mark.skipif(not _have_diffusers() and not _DIFF_FIXTURE.exists(),
                                     reason="diffusers is not installed and tests/golden/diffusers_tiny.npz has not been generated "
                                            "(python tests/golden/make_golden_from_diffusers.py where diffusers exists): parity unpinned")


def _upstream_case():
    """(state_dict, hidden, enc, timestep, velocity, rope_img, rope_txt, tables) from a live diffusers run, else the fixture"""
    sys.path.insert(0, str(GOLD))
    import make_golden_from_diffusers as G
    if _have_diffusers():
        m = G.build_diffusers_tiny()
        hidden, enc, ts, v, rope = G.diffusers_outputs(m)
        sd = {k: t.clone() for k, t in m.state_dict().items()}
        return (sd, hidden, enc, ts, v, torch.view_as_real(rope[0]), torch.view_as_real(rope[1]),
                {k: torch.from_numpy(np.asarray(a)) for k, a in G.scheduler_tables().items()})
    z = np.load(_DIFF_FIXTURE)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    t = lambda k: torch.from_numpy(z[k])
    tables = {k: t(k) for k in z.files if k.startswith(("sigmas_", "timesteps_", "step_"))}
    return sd, t("hidden"), t("enc"), t("timestep"), t("velocity"), t("rope_img"), t("rope_txt"), tables


@_needs_upstream
def test_oracle_matches_diffusers_transformer_on_identical_weights():
    """The oracle loads the diffusers state_dict unchanged (same parameter names) and must reproduce the upstream
    velocity and RoPE tables in fp32 — the check that turns 'parity unpinned' into pinned."""
    sd, hidden, enc, ts, v, rope_img, rope_txt, _ = _upstream_case()
    m = R.QwenImageTransformer2DModelRef(R.TINY_CONFIG).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "pos_embed" not in k] and not [k for k in unexpected if "pos_embed" not in k], (missing, unexpected)
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    with torch.no_grad():
        got = m(hidden, enc, None, ts, shapes, [enc.shape[1]])[0]
        fi, ft = m.pos_embed(shapes, [enc.shape[1]])
    assert torch.allclose(got, v, atol=2e-5, rtol=1e-4), float((got - v).abs().max())
    assert torch.allclose(torch.view_as_real(fi), rope_img, atol=1e-6) and torch.allclose(torch.view_as_real(ft), rope_txt, atol=1e-6)


@_needs_upstream
def test_oracle_matches_diffusers_scheduler():
    *_, tables = _upstream_case()
    for n in (2, 4, 8):
        for seq in (4096, 1024, 256):
            ref = R.ref_flowmatch_sigmas(n, seq)
            assert np.allclose(ref, tables[f"sigmas_n{n}_s{seq}"].numpy(), atol=2e-6), (n, seq)
            assert np.allclose(ref[:-1] * 1000, tables[f"timesteps_n{n}_s{seq}"].numpy(), atol=2e-3), (n, seq)
            assert np.allclose(qie_b200.flowmatch_sigmas(n, seq), tables[f"sigmas_n{n}_s{seq}"].numpy(), atol=2e-6)
    sig = R.ref_flowmatch_sigmas(4, 4096)
    out = R.ref_euler_step(tables["step_x"], tables["step_v"], float(sig[0]), float(sig[1]))
    assert torch.allclose(out, tables["step_out"], atol=1e-6)
