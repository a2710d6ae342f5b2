"""End-to-end parity (-m gpu): the CUDA transformer (through qie_forward) and the fused CFG+Euler loop against the
fp32 oracle on identical bf16-rounded random-init weights, latents and cached embeddings.

Tolerances are BASELINE.json's: per-step velocity max-rel-err = max|a-b| / max|b| <= 2e-2, final-latent cosine >= 0.999.
"""
import dataclasses

import pytest
import torch

import kernels as K
import qie_b200
from oracle import qwen_mmdit_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VEL_TOL = 2e-2
COS_TOL = 0.999


def small_cfg(layers=2, heads=2, joint=128):
    ref = R.RefConfig(num_layers=layers, attention_head_dim=128, num_attention_heads=heads, joint_attention_dim=joint)
    ours = qie_b200.QwenImageDiTConfig(num_layers=layers, num_attention_heads=heads, joint_attention_dim=joint)
    return ref, ours


def build_pair(ref_cfg, our_cfg, seed=0):
    oracle = R.init_weights_(R.QwenImageTransformer2DModelRef(ref_cfg), seed=seed)
    with torch.no_grad():
        for p in oracle.parameters():          # both sides see the same bf16-representable weights
            p.copy_(p.to(torch.bfloat16).float())
    oracle.eval()
    ours = qie_b200.B200QwenImageTransformer2DModel.from_state_dict(oracle.state_dict(), our_cfg, DEV)
    return oracle, ours


def bf16_round(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("fuse_qk", [1, 0])
@pytest.mark.parametrize("B,shapes,T", [(1, [(1, 16, 16), (1, 16, 16)], 19), (2, [(1, 10, 12), (1, 16, 8)], 130)])
def test_single_step_velocity(B, shapes, T, fuse_qk):
    ref_cfg, our_cfg = small_cfg()
    oracle, ours = build_pair(ref_cfg, our_cfg)
    ours.set_option(0, fuse_qk)
    img_shapes = [shapes] * B
    hidden, enc = R.make_inputs(ref_cfg, img_shapes, T, batch=B, seed=1)
    hidden, enc = bf16_round(hidden), bf16_round(enc)
    ts = torch.tensor([0.76953125] * B)
    with torch.no_grad():
        ref = oracle(hidden, enc, None, ts, img_shapes, [T] * B)[0]
        got = ours(hidden_states=hidden.to(DEV), encoder_hidden_states=enc.to(DEV), timestep=ts.to(DEV),
                   img_shapes=img_shapes, txt_seq_lens=[T] * B, return_dict=False)[0]
    assert got.shape == ref.shape
    err = K.rel_err(got.cpu(), ref)
    assert err <= VEL_TOL, err


@pytest.mark.parametrize("shapes,lens", [
    ([(1, 6, 10), (1, 4, 4)], [1]),                              # 76 image tokens (less than one tile), ONE text token
    ([(1, 16, 16), (1, 16, 16)], [1100]),                        # long prompt: 9 text tiles, 2 image tiles... text outweighs the image
    ([(1, 8, 16)], [1, 5, 128, 129, 200, 64, 33, 7]),            # maximum batch (8), every element its own text length
    ([(1, 16, 8), (1, 16, 8), (1, 16, 8), (1, 16, 8)], [255, 257]),   # four reference images; lengths either side of a tile edge
])
def test_edge_shapes_against_oracle(shapes, lens):
    """Edge cases of the operator surface: tiny and ragged inputs, the maximum batch, text longer than the image,
    per-element prompt lengths straddling a 128-row tile boundary — each batch element against its own fp32 oracle forward."""
    ref_cfg, our_cfg = small_cfg()
    oracle, ours = build_pair(ref_cfg, our_cfg)
    B, T = len(lens), max(lens)
    img_shapes = [shapes] * B
    hidden, enc = R.make_inputs(ref_cfg, img_shapes, T, batch=B, seed=4)
    hidden, enc = bf16_round(hidden), bf16_round(enc)
    ts = torch.tensor([0.5] * B)
    got = ours(hidden_states=hidden.to(DEV), encoder_hidden_states=enc.to(DEV), timestep=ts.to(DEV), img_shapes=img_shapes,
               txt_seq_lens=lens, return_dict=False)[0].cpu()
    assert torch.isfinite(got.float()).all()
    with torch.no_grad():
        for b in range(B):
            ref = oracle(hidden[b:b + 1], enc[b:b + 1, :lens[b]], None, ts[b:b + 1], [shapes], [lens[b]])[0]
            err = K.rel_err(got[b:b + 1], ref)
            assert err <= VEL_TOL, (b, lens[b], err)


def _scaled_norm_weights(oracle, factor):
    with torch.no_grad():
        for name, p in oracle.named_parameters():
            if p.dim() == 1 and name.endswith("weight") and any(k in name for k in ("norm_q", "norm_k", "norm_added_q", "norm_added_k")):
                p.mul_(factor)
                p.copy_(p.to(torch.bfloat16).float())


@pytest.mark.parametrize("factor,expect_bounded", [(1.0, True), (2.0, True), (2.6, False), (6.0, False)])
def test_bounded_score_attention_follows_the_norm_weights(factor, expect_bounded):
    """qie_set_weights derives a bound on |q.k| from the QK-RMSNorm weights of every block; blocks under QIE_ATTN_SCORE_BOUND (80 in
    the log2 domain) run the bounded-score attention (q pre-scaled in the QKV epilogue, no running max), the others the
    online-softmax kernel.  Either way the velocity equals the fp32 oracle's, and switching the form off (option 3) agrees."""
    ref_cfg, our_cfg = small_cfg()
    oracle = R.init_weights_(R.QwenImageTransformer2DModelRef(ref_cfg), seed=0)
    with torch.no_grad():
        for p in oracle.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    _scaled_norm_weights(oracle, factor)
    oracle.eval()
    ours = qie_b200.B200QwenImageTransformer2DModel.from_state_dict(oracle.state_dict(), our_cfg, DEV)
    lib = qie_b200.lib()
    sd = oracle.state_dict()
    for l in range(ref_cfg.num_layers):
        names = [k for k in sd if k.startswith(f"transformer_blocks.{l}.") and k.endswith("weight") and
                 any(t in k for t in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"))]
        pair_max = lambda ks: torch.stack([sd[k].abs().view(64, 2).amax(dim=1) for k in ks]).amax(dim=0)      # per RoPE pair, both streams
        mq, mk = pair_max([k for k in names if "_q" in k]), pair_max([k for k in names if "_k" in k])
        want = 128 * (mq * mk).max().item() * 0.08838834764831845 * 1.4426950408889634 * 1.02
        got_b = lib.qie_attn_score_bound(ours._handle, l)
        assert abs(got_b - want) <= 1e-3 * want, (l, got_b, want)
        assert (got_b <= 80.0) == expect_bounded
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    T = 77
    hidden, enc = R.make_inputs(ref_cfg, shapes, T, seed=5)
    hidden, enc = bf16_round(hidden), bf16_round(enc)
    ts = torch.tensor([0.625])
    with torch.no_grad():
        ref = oracle(hidden, enc, None, ts, shapes, [T])[0]
    run = lambda: ours(hidden_states=hidden.to(DEV), encoder_hidden_states=enc.to(DEV), timestep=ts.to(DEV), img_shapes=shapes,
                       txt_seq_lens=[T], return_dict=False)[0].cpu()
    got = run()
    assert K.rel_err(got, ref) <= VEL_TOL, K.rel_err(got, ref)
    ours.set_option(3, 0)
    off = run()
    assert K.rel_err(off, ref) <= VEL_TOL
    assert K.rel_err(got, off.float()) <= VEL_TOL
    if not expect_bounded:
        assert torch.equal(got, off)          # the same kernels ran both times


def test_score_bound_is_taken_per_rope_pair():
    """An outlier norm weight on one channel of q and on ANOTHER channel of k: the product of the maxima (16) would rule the bounded
    form out, the per-pair bound (RoPE only rotates inside a pair) keeps it — and the velocity still equals the oracle's."""
    ref_cfg, our_cfg = small_cfg()
    oracle = R.init_weights_(R.QwenImageTransformer2DModelRef(ref_cfg), seed=0)
    with torch.no_grad():
        for p in oracle.parameters():
            p.copy_(p.to(torch.bfloat16).float())
        for blk in oracle.transformer_blocks:
            blk.attn.norm_q.weight[10] = 4.0
            blk.attn.norm_added_q.weight[11] = 4.0          # same RoPE pair as channel 10
            blk.attn.norm_k.weight[77] = 4.0
            blk.attn.norm_added_k.weight[100] = 4.0
    oracle.eval()
    ours = qie_b200.B200QwenImageTransformer2DModel.from_state_dict(oracle.state_dict(), our_cfg, DEV)
    lib = qie_b200.lib()
    for l in range(ref_cfg.num_layers):
        b = lib.qie_attn_score_bound(ours._handle, l)
        assert 60.0 < b <= 80.0, b                          # 4 x ~1.06 x 16.32 x 1.02; the product of the maxima would give ~270
        assert lib.qie_attn_layer_variant(ours._handle, l) & 0x200
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    T = 40
    hidden, enc = R.make_inputs(ref_cfg, shapes, T, seed=6)
    hidden, enc = bf16_round(hidden), bf16_round(enc)
    ts = torch.tensor([0.375])
    with torch.no_grad():
        ref = oracle(hidden, enc, None, ts, shapes, [T])[0]
    got = ours(hidden_states=hidden.to(DEV), encoder_hidden_states=enc.to(DEV), timestep=ts.to(DEV), img_shapes=shapes,
               txt_seq_lens=[T], return_dict=False)[0].cpu()
    assert K.rel_err(got, ref) <= VEL_TOL, K.rel_err(got, ref)


def test_batch_above_eight_and_bad_text_lengths_are_refused():
    ref_cfg, our_cfg = small_cfg()
    _, ours = build_pair(ref_cfg, our_cfg)
    shapes = [(1, 8, 16)]
    x = torch.zeros(9, 128, 64, dtype=torch.bfloat16, device=DEV)
    e = torch.zeros(9, 8, 128, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(qie_b200.QieError):
        ours(x, e, None, torch.zeros(9, device=DEV), [shapes] * 9, [8] * 9, return_dict=False)
    with pytest.raises(qie_b200.QieError):          # a batch element without text
        ours(x[:2], e[:2], None, torch.zeros(2, device=DEV), [shapes] * 2, [8, 0], return_dict=False)
    with pytest.raises(qie_b200.QieError):          # lengths longer than the embeddings
        ours(x[:2], e[:2], None, torch.zeros(2, device=DEV), [shapes] * 2, [8, 9], return_dict=False)


def test_output_object_and_dtype_surface():
    ref_cfg, our_cfg = small_cfg(layers=1)
    _, ours = build_pair(ref_cfg, our_cfg)
    shapes = [[(1, 8, 16)]]
    hidden, enc = R.make_inputs(ref_cfg, shapes, 7)
    out = ours(hidden.to(DEV).bfloat16(), enc.to(DEV).bfloat16(), None, torch.tensor([1.0], device=DEV).bfloat16(),
               shapes, [7])
    assert out.sample.shape == (1, 128, 64) and out.sample.dtype == torch.bfloat16
    assert ours.config.in_channels == 64 and ours.config.guidance_embeds is False
    assert ours.dtype == torch.bfloat16 and ours.device == torch.device(DEV)
    with ours.cache_context("cond"):
        pass
    assert next(ours.parameters()).device == torch.device(DEV)
    # pipeline.to("cuda") / .eval() reach the module (server.py:70-71): no-ops here; a move off the device or to another dtype
    # must fail loudly (there is no CPU path and the weights are packed bf16)
    assert ours.to("cuda") is ours and ours.to(DEV) is ours and ours.to(torch.bfloat16) is ours and ours.eval() is ours
    for bad in ("cpu", torch.float32):
        with pytest.raises(qie_b200.QieError):
            ours.to(bad)


def test_truncated_stack_matches_oracle():
    ref_cfg, our_cfg = small_cfg(layers=3)
    oracle, ours = build_pair(ref_cfg, our_cfg)
    shapes = [[(1, 16, 16)]]
    hidden, enc = R.make_inputs(ref_cfg, shapes, 40)
    hidden, enc = bf16_round(hidden), bf16_round(enc)
    ts = torch.tensor([0.5])
    with torch.no_grad():
        ref = oracle(hidden, enc, None, ts, shapes, [40], num_blocks=1)[0]
    got = ours(hidden.to(DEV), enc.to(DEV), None, ts.to(DEV), shapes, [40], return_dict=False, num_blocks=1)[0]
    assert K.rel_err(got.cpu(), ref) <= VEL_TOL


@pytest.mark.parametrize("cfg_on", [False, True])
def test_denoise_loop_parity(cfg_on):
    """4-step schedule, per-step velocity and final-latent cosine (configs 2/3 of BASELINE.json at reduced width)."""
    ref_cfg, our_cfg = small_cfg(layers=4)
    oracle, ours = build_pair(ref_cfg, our_cfg)
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    g = torch.Generator().manual_seed(5)
    lat = bf16_round(torch.randn(1, 256, 64, generator=g))
    img_lat = bf16_round(torch.randn(1, 256, 64, generator=g))
    cond = bf16_round(torch.randn(1, 37, 128, generator=g) * 3)
    unc = bf16_round(torch.randn(1, 20, 128, generator=g) * 3) if cfg_on else None
    ref_v = []
    with torch.no_grad():
        ref_final = R.ref_run_denoise(oracle, lat, img_lat, cond, shapes, 4, unc, 4.0, collect=ref_v)
    got_v = []
    got_final = qie_b200.run_denoise(ours, lat.to(DEV), img_lat.to(DEV), cond.to(DEV), shapes, 4,
                                     None if unc is None else unc.to(DEV), 4.0, collect=got_v)
    # per-step velocity of the FIRST step sees identical inputs on both sides
    vc0 = got_v[0][0].float().cpu()
    if cfg_on:
        vu0 = got_v[0][1].float().cpu()
        v0 = R.ref_cfg_combine(vc0, vu0, 4.0)
    else:
        v0 = vc0
    assert K.rel_err(v0, ref_v[0]) <= 2 * VEL_TOL if cfg_on else K.rel_err(v0, ref_v[0]) <= VEL_TOL
    cos = torch.nn.functional.cosine_similarity(got_final.float().cpu().flatten(), ref_final.float().flatten(), dim=0)
    assert cos >= COS_TOL, cos


def test_two_handles_on_two_streams_do_not_share_scratch():
    """the split-K tail scratch / tickets and the adaLN row counters used to be one set per device and
    process, so two forwards on two CUDA streams (cond / uncond on two streams, README.md:127-128) raced on them.  They are now
    one set per (device, stream): two handles driven concurrently from two streams must reproduce their serial results bit for
    bit, repeatedly.  The shapes are chosen so that both mutable paths are live (a K-split tail in FF-down: K = 4 D >= 6144 needs
    D >= 1536 -> 12 heads; the streaming adaLN kernel always takes rows from its counter)."""
    cfg = qie_b200.QwenImageDiTConfig(num_layers=2, num_attention_heads=12, joint_attention_dim=128)
    models = [qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=s_, device=DEV) for s_ in (0, 1)]
    shapes = [[(1, 64, 64)]]      # 16 + 1 m-units x 6 n-blocks = 102 pair-tiles on 74 TPCs: a 28-tile tail, split in two K ranges
    g = torch.Generator(device=DEV).manual_seed(3)
    xs = [torch.randn(1, 4096, 64, generator=g, device=DEV).bfloat16() for _ in range(2)]
    encs = [(torch.randn(1, 256, 128, generator=g, device=DEV) * 3).bfloat16() for _ in range(2)]
    ts = torch.tensor([0.5], device=DEV)
    serial = [m(x, e, None, ts, shapes, [256], return_dict=False)[0].clone() for m, x, e in zip(models, xs, encs)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):
        outs = [None, None]
        for i in (0, 1):
            streams[i].wait_stream(torch.cuda.current_stream())
        for rep in range(2):                      # interleave the enqueueing so that the two forwards really overlap on the device
            for i in (0, 1):
                with torch.cuda.stream(streams[i]):
                    outs[i] = models[i](xs[i], encs[i], None, ts, shapes, [256], return_dict=False)[0]
        torch.cuda.synchronize()
        assert torch.equal(outs[0], serial[0]) and torch.equal(outs[1], serial[1])


def test_batched_cfg_forward_equals_separate_forwards():
    """Batched true CFG (the reference's batched_cfg_pipeline.py, README.md:126): the cond and the uncond prompt — different
    lengths, 37 and 20 tokens here, another pair that pads to different tile counts below — share ONE forward of batch 2, each
    batch element with its own text length (qie_seq::txt_rows_b).  Rows behind an element's length are padding for every kernel
    (masked keys in attention), so the velocities must equal the two separate batch-1 forwards, and the denoise loop built on it
    must equal the unbatched loop and the fp32 oracle."""
    ref_cfg, our_cfg = small_cfg(layers=3)
    oracle, ours = build_pair(ref_cfg, our_cfg)
    shapes = [[(1, 16, 16), (1, 12, 10)]]
    g = torch.Generator().manual_seed(9)
    x = bf16_round(torch.randn(1, 376, 64, generator=g)).to(DEV).bfloat16()
    ts = torch.tensor([0.5], device=DEV)
    for Tc, Tu in ((37, 20), (300, 40), (20, 129)):
        cond = bf16_round(torch.randn(1, Tc, 128, generator=g) * 3).to(DEV).bfloat16()
        unc = bf16_round(torch.randn(1, Tu, 128, generator=g) * 3).to(DEV).bfloat16()
        vc = ours(x, cond, None, ts, shapes, [Tc], return_dict=False)[0]
        vu = ours(x, unc, None, ts, shapes, [Tu], return_dict=False)[0]
        both = torch.zeros(2, max(Tc, Tu), 128, dtype=torch.bfloat16, device=DEV)
        both[0, :Tc], both[1, :Tu] = cond[0], unc[0]
        both[0, Tc:], both[1, Tu:] = 7.0, -9.0          # whatever sits behind an element's length must not matter
        v2 = ours(torch.cat([x, x], 0), both, None, torch.cat([ts, ts], 0), shapes * 2, [Tc, Tu], return_dict=False)[0]
        assert K.rel_err(v2[0], vc[0]) <= 1e-3 and K.rel_err(v2[1], vu[0]) <= 1e-3, (Tc, Tu, K.rel_err(v2[0], vc[0]), K.rel_err(v2[1], vu[0]))
    lat = bf16_round(torch.randn(1, 256, 64, generator=g))
    img_lat = bf16_round(torch.randn(1, 120, 64, generator=g))
    cond = bf16_round(torch.randn(1, 37, 128, generator=g) * 3)
    unc = bf16_round(torch.randn(1, 20, 128, generator=g) * 3)
    plain = qie_b200.run_denoise(ours, lat.to(DEV), img_lat.to(DEV), cond.to(DEV), shapes, 4, unc.to(DEV), 4.0)
    batched = qie_b200.run_denoise(ours, lat.to(DEV), img_lat.to(DEV), cond.to(DEV), shapes, 4, unc.to(DEV), 4.0, batched_cfg=True)
    assert K.rel_err(batched, plain) <= 2e-3
    with torch.no_grad():
        ref = R.ref_run_denoise(oracle, lat, img_lat, cond, shapes, 4, unc, 4.0)
    cos = torch.nn.functional.cosine_similarity(batched.float().cpu().flatten(), ref.flatten(), dim=0)
    assert cos >= COS_TOL, cos


def _config4_gate(oracle, ours, lat, img_lat, cond, shapes, steps, tag, dev_oracle):
    """Config 4: the W8A8 CUDA paths against 'the reference's own int8 path' at MODEL level.  The fp32 oracle and
    the int8 oracle model (the same blocks with ref_int8_linear in the four per-block linear groups, R.quantized_view) denoise
    the same inputs; the CUDA int8 / fp8 paths must not be further from the fp32 oracle than the int8 oracle is:
        err(CUDA int8 vs fp32 oracle) <= 1.05 * err(int8 oracle vs fp32 oracle) + err(CUDA bf16 vs fp32 oracle)
    (relative Frobenius error of the step-0 velocity and of the final latents; the last term is the 16-bit storage of everything
    OUTSIDE the quantised linears, which the fp32 int8-oracle does not pay and which is bounded by the bf16 tolerance of the
    north star).  e4m3 has 3 mantissa bits against int8's 7: on Gaussian-like operands its product error is ~3.4x the int8 one
    in the oracle itself (ref_fp8_linear vs ref_int8_linear), so the FP8 path is held to the FP8 ORACLE by the same rule and its
    ratio to the int8 oracle is reported, not asserted."""
    fro = lambda a, b: ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()
    cosine = lambda a, b: torch.nn.functional.cosine_similarity(a.float().cpu().flatten(), b.float().cpu().flatten(), dim=0).item()
    to_o = lambda t: t.float().to(dev_oracle)
    ref_v, out = [], {}
    with torch.no_grad():
        # the oracles follow the bf16 pipeline's timestep rounding (t = 1000 sigma cast to bf16), everything else in fp32
        ref_final = R.ref_run_denoise(oracle, to_o(lat), to_o(img_lat), to_o(cond), shapes, steps, collect=ref_v,
                                      timestep_dtype=torch.bfloat16)
        for kind in ("int8", "fp8"):
            qv = []
            qf = R.ref_run_denoise(R.quantized_view(oracle, kind), to_o(lat), to_o(img_lat), to_o(cond), shapes, steps, collect=qv,
                                   timestep_dtype=torch.bfloat16)
            out["oracle_" + kind] = (fro(qv[0], ref_v[0]), K.rel_err(qv[0].cpu(), ref_v[0].cpu()), fro(qf, ref_final), cosine(qf, ref_final))
    for mode in ("bf16", "int8", "fp8"):
        ours.set_precision(mode)
        gv = []
        gf = qie_b200.run_denoise(ours, lat.to(DEV), img_lat.to(DEV), cond.to(DEV), shapes, steps, collect=gv)
        out["cuda_" + mode] = (fro(gv[0][0], ref_v[0]), K.rel_err(gv[0][0].cpu(), ref_v[0].cpu()), fro(gf, ref_final), cosine(gf, ref_final))
    ours.set_precision("bf16")
    for k, (e0, m0, ef, c) in out.items():
        print(f"config4[{tag}] {k:12s} step-0 velocity rel-Frobenius {e0:.3e} max-rel {m0:.3e} | final latents rel-Frobenius {ef:.3e} cosine {c:.6f}")
    print(f"config4[{tag}] fp8 / int8 oracle error ratio: {out['oracle_fp8'][0] / out['oracle_int8'][0]:.2f} (oracle), "
          f"{out['cuda_fp8'][0] / out['oracle_int8'][0]:.2f} (CUDA fp8 vs int8 oracle)")
    floor0, floorf = out["cuda_bf16"][0], out["cuda_bf16"][2]
    for mode in ("int8", "fp8"):
        assert out["cuda_" + mode][0] <= 1.05 * out["oracle_" + mode][0] + floor0, (mode, out)
        assert out["cuda_" + mode][2] <= 1.05 * out["oracle_" + mode][2] + floorf, (mode, out)
    assert out["cuda_int8"][3] >= COS_TOL and out["cuda_fp8"][3] >= 0.995, out      # final-latent cosine of the W8A8 paths
    return out


def test_config4_w8a8_paths_against_int8_oracle_model():
    """BASELINE.json configs[3] at reduced depth (2 blocks, 2 heads x 128), 4 steps."""
    ref_cfg, our_cfg = small_cfg(layers=2)
    oracle, ours = build_pair(ref_cfg, our_cfg)
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    g = torch.Generator().manual_seed(5)
    lat = bf16_round(torch.randn(1, 256, 64, generator=g))
    img_lat = bf16_round(torch.randn(1, 256, 64, generator=g))
    cond = bf16_round(torch.randn(1, 37, 128, generator=g) * 3)
    _config4_gate(oracle, ours, lat, img_lat, cond, shapes, 4, "2 blocks", "cpu")
    # single linear with an outlier channel: the two restated products, and the LLM.int8 variant the snapshot configures
    x = torch.randn(256, 256) * 2
    x[:, 3] *= 30
    w = torch.randn(512, 256) / 16
    exact = x @ w.t()
    e_int8, e_fp8, e_llm = (K.rel_err(f(x, w, None), exact) for f in (R.ref_int8_linear, R.ref_fp8_linear, R.ref_llm_int8_linear))
    print(f"one linear with an outlier channel: int8 {e_int8:.3e} fp8 {e_fp8:.3e} llm.int8(6.0) {e_llm:.3e}")
    assert e_llm <= e_int8 and e_fp8 <= 4 * e_int8 + 1e-3


def test_config4_full_size_w8a8_against_int8_oracle_model():
    """BASELINE.json configs[3] at FULL size: 60 blocks, 8192 + 219 tokens, 4 steps; fp32 oracle and int8 / fp8 oracle models on
    the same GPU from the same bf16-rounded random-init weights."""
    free, total = torch.cuda.mem_get_info()
    if total < 150e9:
        pytest.skip("needs a 180 GB device for the fp32 oracle + the bf16 and 8-bit models")
    cfg = qie_b200.QwenImageDiTConfig()
    ours = qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=0, device=DEV)
    with torch.device(DEV):
        oracle = R.QwenImageTransformer2DModelRef(R.FULL_CONFIG)
    oracle.load_state_dict({k: v for k, v in ours.export_state_dict().items()}, strict=True)
    oracle.eval()
    shapes = [[(1, 64, 64), (1, 64, 64)]]
    g = torch.Generator(device=DEV).manual_seed(1)
    lat = torch.randn(1, 4096, 64, generator=g, device=DEV).bfloat16()
    img_lat = torch.randn(1, 4096, 64, generator=g, device=DEV).bfloat16()
    cond = (torch.randn(1, 219, 3584, generator=g, device=DEV) * 3)
    cond[..., [5, 77, 1000, 3000]] *= 50          # massive-activation channels like Qwen2.5-VL hidden states
    cond = cond.bfloat16()
    _config4_gate(oracle, ours, lat, img_lat, cond, shapes, 4, "60 blocks", DEV)
    del oracle
    torch.cuda.empty_cache()


def test_full_size_model_two_step_denoise_parity():
    """BASELINE.json configs[1] at FULL size: 60 blocks, D=3072, 1024x1024 edit (4096 noise + 4096 reference tokens),
    ragged T=219, 2-step Lightning schedule, against the fp32 oracle run on the same GPU (81.7 GB of fp32 weights) from the
    same bf16-rounded random-init weights."""
    free, total = torch.cuda.mem_get_info()
    if total < 150e9:
        pytest.skip("needs a 180 GB device for the fp32 oracle + the bf16 model")
    cfg = qie_b200.QwenImageDiTConfig()
    ours = qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=0, device=DEV)
    with torch.device(DEV):
        oracle = R.QwenImageTransformer2DModelRef(R.FULL_CONFIG)
    oracle.load_state_dict({k: v for k, v in ours.export_state_dict().items()}, strict=True)
    oracle.eval()
    shapes = [[(1, 64, 64), (1, 64, 64)]]
    g = torch.Generator(device=DEV).manual_seed(1)
    lat = torch.randn(1, 4096, 64, generator=g, device=DEV).bfloat16()
    img_lat = torch.randn(1, 4096, 64, generator=g, device=DEV).bfloat16()
    cond = (torch.randn(1, 219, 3584, generator=g, device=DEV) * 3)
    cond[..., [5, 77, 1000, 3000]] *= 50          # massive-activation channels like Qwen2.5-VL hidden states
    cond = cond.bfloat16()
    ref_v, got_v = [], []
    with torch.no_grad():
        ref_final = R.ref_run_denoise(oracle, lat.float(), img_lat.float(), cond.float(), shapes, 2, collect=ref_v)
    got_final = qie_b200.run_denoise(ours, lat, img_lat, cond, shapes, 2, collect=got_v)
    e0 = K.rel_err(got_v[0][0], ref_v[0])
    cos = torch.nn.functional.cosine_similarity(got_final.float().flatten(), ref_final.float().flatten(), dim=0).item()
    print(f"full-size parity: step-0 velocity max-rel-err {e0:.3e}, final-latent cosine {cos:.6f}")
    assert e0 <= VEL_TOL, e0
    assert cos >= COS_TOL, cos
    del oracle
    torch.cuda.empty_cache()


def test_config5_shapes_batch8_three_images_ragged_text():
    """BASELINE.json configs[4] shapes at FULL width, 4 of the 60 blocks: streaming batch of 8 frames at 512x512 with two
    reference images -> img_shapes [(1,32,32)]*3 per frame (3072 image tokens), ragged T = 427; fp32 oracle on the GPU."""
    ref_cfg = R.RefConfig(num_layers=4)
    cfg = qie_b200.QwenImageDiTConfig(num_layers=4)
    ours = qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=3, device=DEV)
    with torch.device(DEV):
        oracle = R.QwenImageTransformer2DModelRef(ref_cfg)
    oracle.load_state_dict(ours.export_state_dict(), strict=True)
    oracle.eval()
    B, T = 8, 427
    shapes = [[(1, 32, 32), (1, 32, 32), (1, 32, 32)]] * B
    g = torch.Generator(device=DEV).manual_seed(11)
    hidden = torch.randn(B, 3072, 64, generator=g, device=DEV).bfloat16()
    enc = (torch.randn(B, T, 3584, generator=g, device=DEV) * 3).bfloat16()
    ts = torch.tensor([1.0, 0.77, 0.46, 0.02, 1.0, 0.77, 0.46, 0.02], device=DEV).bfloat16()   # per-frame timesteps
    with torch.no_grad():
        ref = oracle(hidden.float(), enc.float(), None, ts.float(), shapes, [T] * B)[0]
    got = ours(hidden, enc, None, ts, shapes, [T] * B, return_dict=False)[0]
    for b in range(B):
        assert K.rel_err(got[b], ref[b]) <= VEL_TOL, (b, K.rel_err(got[b], ref[b]))


def test_error_convention_on_device():
    """Status codes + messages through the C ABI on a real device (no exception crosses the boundary)."""
    import ctypes as C
    from qie_b200 import _lib as L
    lib = L.lib()
    bad = L.ModelCfg(2, 2, 64, 64, 64, 128, (C.c_int * 3)(8, 28, 28))             # head_dim 64 is not supported
    h = C.c_void_p()
    assert lib.qie_create(C.byref(bad), 0, C.byref(h)) == -2 and b"head_dim" in lib.qie_last_error()
    ref_cfg, our_cfg = small_cfg(layers=1)
    _, ours = build_pair(ref_cfg, our_cfg)
    shapes = [[(1, 8, 16)]]
    hidden, enc = R.make_inputs(ref_cfg, shapes, 7)
    # a workspace that is too small -> QIE_ENOMEM (-4), reported as QieError by the wrapper
    seq = qie_b200.make_seq(1, 128, 7)
    need = lib.qie_workspace_bytes(ours._handle, C.byref(seq))
    ws = torch.zeros(2048, dtype=torch.uint8, device=DEV)
    out = torch.empty(1, 128, 64, dtype=torch.bfloat16, device=DEV)
    flat = (C.c_int * 3)(1, 8, 16)
    ts = torch.ones(1, device=DEV)
    rc = lib.qie_forward(ours._handle, L.ptr(hidden.to(DEV).bfloat16()), L.ptr(enc.to(DEV).bfloat16()), L.ptr(ts), flat, 1,
                         C.byref(seq), L.ptr(out), C.c_void_p((ws.data_ptr() + 1023) // 1024 * 1024), 1024, -1, L.cur_stream())
    assert rc == -4 and need > 1024 and b"workspace" in lib.qie_last_error()
    # img_shapes that do not cover the tokens -> QIE_ESHAPE through the python wrapper
    with pytest.raises(qie_b200.QieError, match="img_shapes"):
        ours(hidden.to(DEV), enc.to(DEV), None, ts, [[(1, 8, 8)]], [7])
    # weights not set
    bare = qie_b200.B200QwenImageTransformer2DModel(our_cfg, DEV)
    with pytest.raises(qie_b200.QieError, match="weights not set"):
        bare(hidden.to(DEV), enc.to(DEV), None, ts, shapes, [7])


def test_schedule_and_prompt_caches_are_bit_identical():
    """Next-row N1: cached per-timestep modulation tables and cached per-prompt text embeddings must not
    change a single bit of the denoise result."""
    ref_cfg, our_cfg = small_cfg(layers=3)
    _, ours = build_pair(ref_cfg, our_cfg)
    shapes = [[(1, 16, 16), (1, 12, 10)]]
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(1, 256, 64, generator=g).bfloat16().to(DEV)
    img_lat = torch.randn(1, 120, 64, generator=g).bfloat16().to(DEV)
    cond = (torch.randn(1, 37, 128, generator=g) * 3).bfloat16().to(DEV)
    unc = (torch.randn(1, 22, 128, generator=g) * 3).bfloat16().to(DEV)
    plain = qie_b200.run_denoise(ours, lat, img_lat, cond, shapes, 4, unc, 4.0)
    sig = qie_b200.flowmatch_sigmas(4, 256)
    ours.cache_schedule([float(qie_b200.model_timestep(float(s), 1, "cpu")[0]) for s in sig[:4]])
    ours.cache_prompt("cond", cond).cache_prompt("uncond", unc)
    n0 = qie_b200.lib().qie_launch_count()
    cached = qie_b200.run_denoise(ours, lat, img_lat, cond, shapes, 4, unc, 4.0, use_caches=True)
    n_cached = qie_b200.lib().qie_launch_count() - n0
    n0 = qie_b200.lib().qie_launch_count()
    again = qie_b200.run_denoise(ours, lat, img_lat, cond, shapes, 4, unc, 4.0)
    n_plain = qie_b200.lib().qie_launch_count() - n0
    assert torch.equal(plain, cached) and torch.equal(plain, again)
    assert n_cached == n_plain - 8 * 7        # 8 forwards x (timestep proj + 4 gemv + text rmsnorm + txt_in GEMM) skipped


def test_lora_merge_matches_unmerged_side_path():
    """Next-row N2: offline LoRA merge == PEFT's un-merged x@A^T@B^T*scale side path (server.py:76-79)."""
    ref_cfg, our_cfg = small_cfg(layers=2)
    oracle, _ = build_pair(ref_cfg, our_cfg)
    sd = {k: v.clone() for k, v in oracle.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    lora = {}
    for mod in ("transformer_blocks.0.attn.to_q", "transformer_blocks.1.img_mlp.net.2", "transformer_blocks.1.attn.add_k_proj"):
        w = sd[mod + ".weight"]
        lora["transformer." + mod + ".lora_A.weight"] = torch.randn(8, w.shape[1], generator=g) * 0.05
        lora["transformer." + mod + ".lora_B.weight"] = torch.randn(w.shape[0], 8, generator=g) * 0.05
        lora["transformer." + mod + ".alpha"] = torch.tensor(4.0)
    merged = qie_b200.merge_lora(sd, lora, scale=1.0)
    ours = qie_b200.B200QwenImageTransformer2DModel.from_state_dict(merged, our_cfg, DEV)
    # reference: the oracle with the same merged weights rounded to bf16 (what the packed model holds)
    oracle.load_state_dict({k: v.to(torch.bfloat16).float() for k, v in merged.items()})
    w0 = sd["transformer_blocks.0.attn.to_q.weight"]
    expect = w0 + 0.5 * lora["transformer.transformer_blocks.0.attn.to_q.lora_B.weight"] @ lora["transformer.transformer_blocks.0.attn.to_q.lora_A.weight"]
    assert torch.allclose(merged["transformer_blocks.0.attn.to_q.weight"], expect, atol=1e-6)      # alpha / r = 4 / 8
    shapes = [[(1, 16, 16)]]
    hidden, enc = R.make_inputs(ref_cfg, shapes, 19)
    hidden, enc = bf16_round(hidden), bf16_round(enc)
    ts = torch.tensor([0.5])
    with torch.no_grad():
        ref = oracle(hidden, enc, None, ts, shapes, [19])[0]
    got = ours(hidden.to(DEV), enc.to(DEV), None, ts.to(DEV), shapes, [19], return_dict=False)[0]
    assert K.rel_err(got.cpu(), ref) <= VEL_TOL


# ------------------------------------------------------------------ next rows N3 / N4
@pytest.mark.parametrize("B,h,w", [(1, 128, 128), (2, 6, 10), (1, 64, 96)])
def test_pack_unpack_latents_bit_exact(B, h, w):
    """qie_pack_latents / qie_unpack_latents against the restated _pack_latents / _unpack_latents (+ VAE normalisation)."""
    g = torch.Generator().manual_seed(21)
    C_ = 16
    z = torch.randn(B, C_, h, w, generator=g).bfloat16()
    mean, std = torch.randn(C_, generator=g), torch.rand(C_, generator=g) + 0.5
    # pure re-layout: bit exact
    tok = qie_b200.pack_latents(z.to(DEV))
    assert torch.equal(tok.cpu(), R.ref_pack_latents(z))
    assert torch.equal(qie_b200.unpack_latents(tok, h, w).cpu(), z.unsqueeze(2))
    # fused normalisation: fp32 arithmetic on the bf16 inputs, one rounding -> equal to the fp32 oracle rounded once
    tokn = qie_b200.pack_latents(z.to(DEV), mean, std)
    assert torch.equal(tokn.cpu(), R.ref_pack_latents(z.float(), mean, std).bfloat16())
    zn = qie_b200.unpack_latents(tokn, h, w, mean, std)
    assert torch.equal(zn.cpu(), R.ref_unpack_latents(tokn.cpu().float(), h, w, mean, std).bfloat16())
    with pytest.raises(qie_b200.QieError):
        qie_b200.unpack_latents(tok, h + 2, w)


def test_streaming_frames_match_oracle():
    """N4: key frame, streamed frame (enters the schedule at begin_index with prev_latent + 0.05*noise), key frame — the CUDA
    path against the fp32 oracle running the same state machine."""
    ref_cfg = R.RefConfig(num_layers=2, attention_head_dim=128, num_attention_heads=2, joint_attention_dim=128)
    cfg = qie_b200.QwenImageDiTConfig(num_layers=2, num_attention_heads=2, joint_attention_dim=128)
    oracle = R.init_weights_(R.QwenImageTransformer2DModelRef(ref_cfg), seed=0).eval()
    with torch.no_grad():
        for p_ in oracle.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    model = qie_b200.B200QwenImageTransformer2DModel.from_state_dict(oracle.state_dict(), cfg, DEV)
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    g = torch.Generator().manual_seed(33)
    rnd = lambda *s_: torch.randn(*s_, generator=g).to(torch.bfloat16).float()
    cond = rnd(1, 19, 128) * 3
    sd = qie_b200.StreamingDenoiser(model, shapes, cond.to(DEV), num_inference_steps=4, keyframe_interval=2, stream_steps=2)
    prev = None
    for f in range(3):
        img_lat, noise = rnd(1, 256, 64), rnd(1, 256, 64)
        start, key = R.ref_stream_prepare_latent(prev, noise, f, 2, 0.05)
        with torch.no_grad():
            want = R.ref_run_denoise(oracle, start, img_lat, cond, shapes, 4, begin_index=0 if key else 2)
        got = sd.process_frame(img_lat.to(DEV), noise.to(DEV).bfloat16())
        assert sd.is_keyframe == key
        cos = torch.nn.functional.cosine_similarity(got.float().cpu().flatten(), want.flatten(), dim=0).item()
        assert cos >= 0.999, (f, cos)
        prev = want


def test_forward_is_cuda_graph_capturable():
    """include/qie.h: qie_forward enqueues on the caller's stream without allocating or synchronising in the steady state, so
    the whole 2-block step can be captured once and replayed (what a serving loop does to drop the launch overhead)."""
    ref_cfg = R.RefConfig(num_layers=2, attention_head_dim=128, num_attention_heads=2, joint_attention_dim=128)
    cfg = qie_b200.QwenImageDiTConfig(num_layers=2, num_attention_heads=2, joint_attention_dim=128)
    oracle = R.init_weights_(R.QwenImageTransformer2DModelRef(ref_cfg), seed=0)
    model = qie_b200.B200QwenImageTransformer2DModel.from_state_dict(oracle.state_dict(), cfg, DEV)
    shapes = [[(1, 16, 16), (1, 16, 16)]]
    g = torch.Generator().manual_seed(44)
    x = torch.randn(1, 512, 64, generator=g).bfloat16().to(DEV)
    cond = (torch.randn(1, 19, 128, generator=g) * 3).bfloat16().to(DEV)
    ts = torch.tensor([0.5], device=DEV)
    eager = model(x, cond, None, ts, shapes, [19], return_dict=False)[0].clone()      # also builds the RoPE table (one-off sync)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model(x, cond, None, ts, shapes, [19], return_dict=False)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = model(x, cond, None, ts, shapes, [19], return_dict=False)[0]
    x2 = torch.randn(1, 512, 64, generator=g).bfloat16().to(DEV)
    want2 = model(x2, cond, None, ts, shapes, [19], return_dict=False)[0].clone()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    x.copy_(x2)                       # new input in the captured buffer, same graph
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want2)


@pytest.mark.parametrize("precision", ["bf16", "fp8"])
def test_programmatic_dependent_launch_is_bit_identical(precision):
    """qie_tune(7, v): GEMM / attention / adaLN launched with the programmatic-serialisation attribute (prologue under the
    previous kernel's tail, griddepcontrol.wait before the first global access) vs plain stream order — eager, repeated
    (a race would show as run-to-run differences) and captured in a CUDA graph."""
    ref_cfg = R.RefConfig(num_layers=3, attention_head_dim=128, num_attention_heads=2, joint_attention_dim=128)
    cfg = qie_b200.QwenImageDiTConfig(num_layers=3, num_attention_heads=2, joint_attention_dim=128)
    oracle = R.init_weights_(R.QwenImageTransformer2DModelRef(ref_cfg), seed=0)
    model = qie_b200.B200QwenImageTransformer2DModel.from_state_dict(oracle.state_dict(), cfg, DEV)
    if precision != "bf16":
        model.set_precision(precision)
    shapes = [[(1, 32, 32), (1, 24, 20)]] * 2
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 1504, 64, generator=g).bfloat16().to(DEV)
    cond = (torch.randn(2, 37, 128, generator=g) * 3).bfloat16().to(DEV)
    ts = torch.tensor([0.5, 0.5], device=DEV)
    lib = qie_b200.lib()
    prev = lib.qie_tune_get(7)
    try:
        assert lib.qie_tune(7, 0) == 0
        plain = model(x, cond, None, ts, shapes, [37, 37], return_dict=False)[0].clone()
        assert lib.qie_tune(7, 1) == 0
        for _ in range(5):
            got = model(x, cond, None, ts, shapes, [37, 37], return_dict=False)[0]
            assert torch.equal(got, plain)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model(x, cond, None, ts, shapes, [37, 37], return_dict=False)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = model(x, cond, None, ts, shapes, [37, 37], return_dict=False)[0]
        for _ in range(3):
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, plain)
    finally:
        lib.qie_tune(7, prev)


@pytest.mark.slow
def test_long_sequence_256x256_latent_tiny_model():
    """BASELINE.json configs[0] read literally: a 256x256 latent = 128x128 packed tokens of noise plus as many
    reference tokens -> 32768 image tokens + ragged text through the reduced-width 2-block model, single step at sigma = 1.
    The longest sequence the suite runs: 129 KV tiles of 256 rows per head, RoPE h/w indices -64..63."""
    ref_cfg, our_cfg = small_cfg(layers=2)
    oracle, ours = build_pair(ref_cfg, our_cfg)
    oracle = oracle.to(DEV)
    shapes = [[(1, 128, 128), (1, 128, 128)]]
    hidden, enc = R.make_inputs(ref_cfg, shapes, 19, seed=1)
    hidden, enc = bf16_round(hidden).to(DEV), bf16_round(enc).to(DEV)
    ts = torch.tensor([1.0], device=DEV)
    with torch.no_grad():
        ref = oracle(hidden, enc, None, ts, shapes, [19])[0]
    got = ours(hidden_states=hidden, encoder_hidden_states=enc, timestep=ts, img_shapes=shapes, txt_seq_lens=[19],
               return_dict=False)[0]
    assert got.shape == (1, 32768, 64)
    err = K.rel_err(got, ref)
    assert err <= VEL_TOL, err
