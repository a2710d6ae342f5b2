#!/usr/bin/env python
"""Headline benchmark: edited 1024x1024 images / s, 2-step Lightning schedule (BASELINE.json configs[1]).

One "step" = one edited image = the hot path of the reference over one synthetic input:
2 x QwenImageTransformer2DModel forward (60 blocks, 8192 image + 256 text tokens) + 2 x fused CFG/Euler update.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libqie.so)
  python bench.py --impl reference ...                      # the reference's CPU path (the fp32 oracle restatement of
                                                            # the diffusers transformer; diffusers itself is not installable)
Launch with torchrun for N > 1 (one rank per GPU).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "edited_1024x1024_images_per_s_2step_lightning"
UNIT = "img/s"
IMG_SHAPES = [[(1, 64, 64), (1, 64, 64)]]     # noise latents + one 1024^2 reference image
N_NOISE, N_IMG_TOK, T_TXT = 4096, 8192, 256
STEPS_PER_IMAGE = 2


def flops_per_forward(cfg_layers=60, D=3072, H=24, S_i=None, T=None, B=1):
    """per block 2*D*(3D + D + 2*FF)*S linear + 4*S^2*d_h*H attention (+ tiny top)."""
    S = (N_IMG_TOK if S_i is None else S_i) + (T_TXT if T is None else T)
    lin = 2.0 * D * (3 * D + D + 8 * D) * S
    att = 4.0 * S * S * 128 * H
    return B * cfg_layers * (lin + att)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1404.4), d.get("bf16_tflops", 1638.8), d.get("hbm_gbs", 6556.2), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def q8_library_peak(kind: str, device, seconds: float = 2.0):
    """The 8-bit tensor peak the W8A8 GEMMs are held against, measured in this run like MEASURED_PEAKS.json measures the bf16
    one: the library GEMM (cuBLASLt e4m3 through torch._scaled_mm / int8 through torch._int_mm) on 8192^3, best of 10 (burst)
    and back to back for `seconds` (sustained, the figure for a kernel timed inside a long step).  (sustained, burst, how) in
    T(FL)OP/s, or None when this torch build does not offer the entry point."""
    import time
    n = 8192
    try:
        if kind == "fp8":
            a = torch.randn(n, n, device=device).to(torch.float8_e4m3fn)
            b = torch.randn(n, n, device=device).to(torch.float8_e4m3fn).t()
            one = torch.ones((), device=device)
            fn = lambda: torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
            how = "torch._scaled_mm e4m3 (cuBLASLt) 8192^3"
        else:
            a = torch.randint(-127, 128, (n, n), device=device, dtype=torch.int8)
            b = torch.randint(-127, 128, (n, n), device=device, dtype=torch.int8).t()
            fn = lambda: torch._int_mm(a, b)
            how = "torch._int_mm int8 (cuBLASLt) 8192^3"
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(device)
            best = min(best, e0.elapsed_time(e1))
        ms, t_end = [], time.time() + seconds
        while time.time() < t_end:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record(); torch.cuda.synchronize(device)
            ms.append(e0.elapsed_time(e1) / 10)
        half = ms[len(ms) // 2:]
        flops = 2.0 * n ** 3
        return flops / (sum(half) / len(half)) / 1e9, flops / best / 1e9, how
    except Exception as e:          # noqa: BLE001 — reported in peak_source, the line falls back to 2 x the bf16 peak
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows if len(r) >= 7)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the fp32 oracle (restated diffusers transformer) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(max_seconds_per_sample: float = 0.0):
    """Returns (run_sample, scale, description): run_sample() executes ONE full-width transformer block of the oracle in
    fp32 on `tokens` joint tokens; `scale` converts its time into seconds per edited image by FLOP proportion.
    The per-sample budget (default 8 s; QIE_BENCH_CPU_SAMPLE_S overrides it, the CPU tests use a fraction of a second)
    bounds the sequence length of the sample, never its width."""
    max_seconds_per_sample = max_seconds_per_sample or float(os.environ.get("QIE_BENCH_CPU_SAMPLE_S", "8.0"))
    from oracle import qwen_mmdit_ref as R
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = R.FULL_CONFIG
    D = cfg.inner_dim
    blk = R.QwenImageTransformerBlock(D, cfg.num_attention_heads, cfg.attention_head_dim).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.normal_(0, 0.02)
    rope = R.QwenEmbedRope(10000, list(cfg.axes_dims_rope), scale_rope=True)

    def make(n_side, t_txt):
        shapes = [(1, n_side, n_side), (1, n_side, n_side)]
        fr = rope(shapes, [t_txt])
        s_i = 2 * n_side * n_side
        g = torch.Generator().manual_seed(0)
        return (torch.randn(1, s_i, D, generator=g), torch.randn(1, t_txt, D, generator=g),
                torch.randn(1, D, generator=g), fr, s_i)

    def block_flops(s_i, t):
        S = s_i + t
        return 2.0 * D * 12 * D * S + 4.0 * S * S * 128 * cfg.num_attention_heads

    # calibrate on a quarter-size sequence, then pick the largest sample that fits the per-step budget
    h, e, temb, fr, s_i = make(32, 64)
    with torch.no_grad():
        t0 = time.perf_counter()
        blk(h, e, temb, fr)
        t_cal = time.perf_counter() - t0
    rate = block_flops(s_i, 64) / t_cal
    side, t_txt = 64, T_TXT
    while side > 16 and block_flops(2 * side * side, t_txt) / rate > max_seconds_per_sample:
        side //= 2
        t_txt = max(32, t_txt // 2)
    h, e, temb, fr, s_i = make(side, t_txt)
    image_flops = STEPS_PER_IMAGE * flops_per_forward()
    scale = image_flops / block_flops(s_i, t_txt)

    def run_sample():
        with torch.no_grad():
            blk(h, e, temb, fr)

    desc = (f"1 of 60 full-width (D=3072, 24 heads) fp32 oracle blocks on {s_i}+{t_txt} tokens per step; seconds per image "
            f"extrapolated by algorithmic-FLOP ratio x{scale:.1f} (2 forwards x 60 blocks at 8192+256 tokens)")
    return run_sample, scale, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run_sample, scale, desc = cpu_reference_step_factory()
    for _ in range(args.warmup):
        run_sample()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_sample()
    dt = (time.perf_counter() - t0) / args.steps
    sec_per_image = dt * scale
    v = 1.0 / sec_per_image
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_per_image * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_eager_gpu(args, emit=True, min_ms=3000.0):
    """Like-for-like GPU baseline: the kernels the reference's diffusers transformer dispatches on this GPU,
    i.e. the restated module in bf16 PyTorch eager (cuBLAS linears, fused SDPA, elementwise ATen kernels).  `--layers` (default
    4 here) full-width blocks at the headline sequence are timed with CUDA events and scaled to 60 blocks x forwards per image
    by block count; none of this repo's kernels run.  Reported next to the headline, never as it."""
    from oracle import qwen_mmdit_ref as R
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    cfg = R.FULL_CONFIG
    D, n_blk = cfg.inner_dim, args.layers if args.layers < 60 else 4
    blocks = [R.QwenImageTransformerBlock(D, cfg.num_attention_heads, cfg.attention_head_dim).eval() for _ in range(n_blk)]
    with torch.no_grad():
        for b in blocks:
            for q in b.parameters():
                q.normal_(0, 0.02)
            b.to(dev, torch.bfloat16)
    side = 64 if args.workload == "1024x1ref" else 32
    shapes = [(1, side, side)] * (2 if args.workload == "1024x1ref" else 3)
    fr = tuple(f.to(dev) for f in R.QwenEmbedRope(10000, list(cfg.axes_dims_rope), scale_rope=True)(shapes, [T_TXT]))
    g = torch.Generator(device=dev).manual_seed(0)
    h = torch.randn(1, N_IMG_TOK, D, generator=g, device=dev).bfloat16()
    e = torch.randn(1, T_TXT, D, generator=g, device=dev).bfloat16()
    temb = torch.randn(1, D, generator=g, device=dev).bfloat16()

    def step():
        with torch.no_grad():
            x, y = h, e
            for b in blocks:
                y, x = b(x, y, temb, fr)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(args.warmup, 3)):
        step()
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    # a 60-block forward keeps the GPU at its power cap for hundreds of ms: repeat the sample for >= 3 s (warm, then timed) so
    # that the blocks run at the sustained clock the headline arm sees, not at the burst clock of a 15 ms sample
    reps = max(args.steps, int(min_ms / max(e0.elapsed_time(e1), 1e-3)) + 1)
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    with ClockSampler(dev.index or 0) as clk:
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
    ms_blk = e0.elapsed_time(e1) / reps / n_blk
    fwd = STEPS_PER_IMAGE * (2 if args.cfg else 1)
    ms_img = ms_blk * 60 * fwd
    line = {"impl": "eager_gpu_oracle", "metric": METRIC, "value": 1e3 / ms_img, "unit": UNIT, "n_gpus": 1,
            "steps": reps, "warmup": reps, "ms_per_step": ms_img, "higher_is_better": True,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args, 1), "clocks": clk.summary(),
            "sample": f"{n_blk} of 60 full-width blocks at {N_IMG_TOK}+{T_TXT} tokens in bf16 PyTorch eager "
                      f"(torch {torch.__version__}: cuBLAS + fused SDPA), repeated {reps}x back to back (>= {min_ms / 1e3:.0f} s at the power cap), x{60 // n_blk} "
                      "by block count; the top "
                      "(embeddings, norm_out, proj_out: < 0.1 % of the FLOPs) is not included",
            "ms_per_block": ms_blk, "dit_forward_ms": ms_blk * 60,
            "step_tflops": fwd * flops_per_forward(60) / (ms_img * 1e-3) / 1e12, "gpu_launches": 0}
    del blocks, h, e, temb
    torch.cuda.empty_cache()
    if emit:
        print(json.dumps(line), flush=True)
    return line


def workload_config(args, n):
    what = (f"1024x1024 single-image edit, {STEPS_PER_IMAGE}-step " + ("Lightning " if STEPS_PER_IMAGE == 2 else "") + "schedule, " if getattr(args, "workload", "1024x1ref") == "1024x1ref"
            else "512x512 frame with two reference images (one frame of BASELINE configs[4]), 4-step schedule, ")
    return {"workload": "Qwen-Image-Edit-2509 MMDiT denoise (60 blocks, D=3072, 24 heads, random-init), " + what +
                        ("true-CFG 4.0 (cond+uncond)" if args.cfg else "cond-only"),
            "img_tokens": N_IMG_TOK, "txt_tokens": T_TXT, "forwards_per_image": STEPS_PER_IMAGE * (2 if args.cfg else 1),
            "frames_per_forward": getattr(args, "batch", 1),
            "layers": args.layers, "precision": args.precision, "caches": bool(getattr(args, "cache", False)),
            "parallelism": (f"dp{n}: one independent frame stream per GPU, weights replicated, no data-path collective"
                            if args.mode == "dp" else f"{args.mode}{' (fused peer-memory exchange)' if args.fused else ' (NCCL all-to-all)' if 'ulysses' in args.mode else ''} over {n} GPUs: ONE frame, weights replicated"),
            "l2": "inputs larger than L2: 40.9 GB of weights + 0.6 GB of activations stream through the 126 MB L2 every forward"}


# ------------------------------------------------------------------------------------------------
# strong-scaling leg: ONE true-CFG frame spread over all ranks (the partition north_star names), checked in-process
# ------------------------------------------------------------------------------------------------
def strong_mode(world):
    if world == 1:
        return "single", 1, 1
    if world == 2:
        return "cfgpair", 2, 1
    if world % 2 == 0 and 24 % (world // 2) == 0:
        return f"cfgpair x fused-ulysses{world // 2}", 2, world // 2
    if 24 % world == 0:
        return f"fused-ulysses{world}", 1, world
    return None, 0, 0


def strong_check(rec: dict, sp: int) -> dict:
    """Verdict on a multi-GPU `strong` record (sp = 1: CFG pair only, > 1: sequence-parallel ranks per branch).
    Two levels.  `parity_tolerance` (1e-2 for the sequence-parallel modes, bit-identical for the CFG pair) is what the partition
    is held to and `parity_ok` says whether it was met.  The run FAILS (exit 1) on a wrong partition: a CFG pair that is not
    bit-identical, a barrier time-out, NaN, a final-latent cosine under 0.999, or a step-0 velocity beyond the north star's own
    bound for a bf16 forward against the reference (2e-2) — the sequence-parallel forward differs from the single-GPU one only by
    the fp32 summation order of the GEMM tail tiles (a few bf16 ulps after 60 blocks: 7.0e-3 on every build measured so far),
    so a value between the two is reported, not fatal."""
    rec["parity_ok"] = bool(rec["parity_err"] <= rec["parity_tolerance"])
    rec["parity_hard_limit"] = 0.0 if sp == 1 else 2e-2
    bad = (rec["parity_err"] > rec["parity_hard_limit"] or rec["barrier_timeouts"] != 0 or not (rec["parity_err"] == rec["parity_err"])
           or rec["final_latent_cosine"] < 0.999 or (sp == 1 and rec["final_latent_max_rel_err"] != 0.0))
    if bad:
        rec["failed"] = "parity or barrier check failed"
    return rec


def strong_leg(args, model, dev, world, rank, timed):
    """ONE edited frame with true CFG (2 steps x (cond + uncond) = 4 forwards) over all N GPUs: CFG pair at N = 2, CFG pair x
    fused (peer-memory) Ulysses at N = 4 / 8.  Every rank first runs the same frame alone (the N = 1 time measured on this box
    in this run, and the reference result), then the group runs it together; the final latents must agree (CFG pair: bit for
    bit; sequence parallel: <= 1e-2 max-rel, only the attention reduction order differs).  Returns the `strong` record; raises
    SystemExit(1) after printing it when the partition is wrong."""
    import torch.distributed as dist
    import qie_b200
    mode, branches, sp = strong_mode(world)
    if mode is None:
        return {"mode": None, "skipped": f"{world} ranks do not factor into cfg branches x a divisor of 24 heads"}
    cfg = model.cfg
    g = torch.Generator(device=dev).manual_seed(1)            # one frame: identical inputs on every rank
    lat = torch.randn(1, N_NOISE, 64, generator=g, device=dev).bfloat16()
    img_lat = torch.randn(1, N_IMG_TOK - N_NOISE, 64, generator=g, device=dev).bfloat16()
    cond = (torch.randn(1, T_TXT, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16()
    unc = (torch.randn(1, T_TXT, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16()

    def single(collect=None):
        return qie_b200.run_denoise(model, lat, img_lat, cond, IMG_SHAPES, STEPS_PER_IMAGE, unc, 4.0, collect=collect)

    ref_v = []
    ref = single(ref_v)
    single()
    n1_ms = timed(single, max(2, args.steps // 2)) / max(2, args.steps // 2)      # max over ranks of the same single-GPU frame
    rec = {"mode": mode, "frame": f"true-CFG 4.0, {STEPS_PER_IMAGE} steps = {2 * STEPS_PER_IMAGE} forwards of {N_IMG_TOK}+{T_TXT} tokens",
           "n1_ms_per_frame": n1_ms, "n1_how": "the same frame on ONE GPU of this box (every rank runs it alone; max over ranks)"}
    if world == 1:
        rec.update({"ms_per_frame": n1_ms, "efficiency_vs_n1": 1.0, "parity_err": 0.0, "barrier_timeouts": 0,
                    "exchange_bytes": 0})
        return rec
    layout = qie_b200.make_layout(world, rank, branches)
    runner = model
    if layout.sp_size > 1:
        runner = qie_b200.UlyssesTransformer(model, layout.sp_group, fused=True)

    def frame(collect=None):
        return qie_b200.run_denoise_parallel(runner, layout, lat, img_lat, cond, unc, IMG_SHAPES, STEPS_PER_IMAGE, 4.0, collect=collect)

    for _ in range(3):                  # eager, graph capture, first replay
        got = frame()
    with ClockSampler(dev.index or 0) as sclk:
        ms = timed(frame, max(args.steps, 8)) / max(args.steps, 8)
    got_v = []
    got = frame(got_v)
    torch.cuda.synchronize()
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
    # the forwards of step 0 see identical inputs on both sides: velocity of the cond and of the uncond forward (max-rel, as the
    # north star states the tolerance); then the frame's final latents (after the CFG combine, which amplifies differences ~4x)
    err = max(rel(got_v[0][0], ref_v[0][0]), rel(got_v[0][1], ref_v[0][1]))
    err_final = rel(got, ref)
    cos = torch.nn.functional.cosine_similarity(got.float().flatten(), ref.float().flatten(), dim=0).item()
    errs = torch.tensor([err, err_final, -cos], device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    timeouts = torch.tensor([qie_b200.lib().qie_peer_barrier_timeouts()], device=dev)
    dist.all_reduce(timeouts, op=dist.ReduceOp.MAX)
    fwd = 2 * STEPS_PER_IMAGE // branches        # forwards each rank takes part in
    xbytes = 0
    prof = None
    if layout.sp_size > 1:
        plan = qie_b200.make_shard_plan(N_IMG_TOK, T_TXT, layout.sp_size, layout.sp_rank)
        xbytes = fwd * qie_b200.exchange_bytes_per_forward(plan, 1, cfg.num_attention_heads, cfg.num_layers)
        # where the time of a sequence-parallel forward goes on this rank: per-class CUDA events of eager launches (no graph)
        model.profile(True)
        frame()
        torch.cuda.synchronize()
        tl = model.read_timeline(8192)
        pr = model.read_profile()
        t0 = time.perf_counter()
        frame()
        torch.cuda.synchronize()
        eager_ms = (time.perf_counter() - t0) * 1e3
        model.read_profile()
        model.profile(False)
        busy = sum(v["ms"] for v in pr.values())
        span = max((a + d for a, d, _ in tl), default=0.0)
        prof = {"per_frame_ms": {k: round(v["ms"], 3) for k, v in pr.items()}, "launches": {k: v["launches"] for k, v in pr.items()},
                "kernel_busy_ms": round(busy, 3), "eager_frame_ms_host_clock": round(eager_ms, 3),
                "first_to_last_launch_ms": round(span, 3),
                "note": "rank 0, one frame launched eagerly with an event pair around every launch (the timed frames replay CUDA "
                        "graphs); barrier = time inside peer_barrier_kernel = waiting for the slowest rank + NVLink flag round trip"}
    if branches == 2:
        xbytes += STEPS_PER_IMAGE * N_NOISE * 64 * 2         # the velocity all-gather of the CFG pair (NCCL, 512 KB per step)
    rec.update({"ms_per_frame": ms, "efficiency_vs_n1": n1_ms / (world * ms), "speedup_vs_n1": n1_ms / ms,
                "parity_err": float(errs[0].item()), "parity_tolerance": 0.0 if sp == 1 else 1e-2,
                "parity_what": "max-rel-err of the step-0 velocities (cond and uncond forward) against the single-GPU forward on the same inputs",
                "final_latent_max_rel_err": float(errs[1].item()), "final_latent_cosine": float(-errs[2].item()),
                "barrier_timeouts": int(timeouts.item()), "exchange_bytes": int(xbytes), "clocks": sclk.summary(),
                "exchange_bytes_note": "bytes ONE rank stores into other GPUs' memory per frame (epilogue peer stores over NVLink "
                                       "+ the CFG-pair velocity all-gather)", "profile": prof})
    if hasattr(runner, "close"):
        runner.close()
    return strong_check(rec, sp)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import qie_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = qie_b200.QwenImageDiTConfig(num_layers=args.layers)
    model = qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=0, device=dev)
    if args.precision != "bf16":
        model.set_precision(args.precision)
    if args.attn_variant:
        model.set_option(1, args.attn_variant)
    for env, key in (("QIE_L2_HINTS", 2), ("QIE_LN_VARIANT", 3), ("QIE_SPLIT_TAIL", 4), ("QIE_GROUP_M", 5), ("QIE_PDL", 7)):   # A/B switches (qie_tune keys)
        if os.environ.get(env):
            qie_b200.lib().qie_tune(key, int(os.environ[env]))
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    NB = args.batch
    lat = torch.randn(NB, N_NOISE, 64, generator=g, device=dev).bfloat16()
    img_lat = torch.randn(NB, N_IMG_TOK - N_NOISE, 64, generator=g, device=dev).bfloat16()
    cond = (torch.randn(NB, T_TXT, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16()
    unc = (torch.randn(NB, T_TXT, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16() if args.cfg else None
    host = {k: v.cpu().pin_memory() for k, v in (("lat", lat), ("img", img_lat), ("cond", cond))}
    if unc is not None:
        host["unc"] = unc.cpu().pin_memory()
    out_host = torch.empty(lat.shape, dtype=torch.bfloat16).pin_memory()

    # --mode dp (default): every rank edits its own frame.  Other modes spread ONE frame over the ranks (strong scaling):
    #   cfgpair: cond / uncond branches on 2 GPU groups;  ulysses: sequence-parallel forward;  cfg+ulysses: both.
    runner, layout = model, None
    if args.mode != "dp":
        if world < 2:
            raise SystemExit("--mode cfgpair/ulysses needs torchrun with >= 2 ranks")
        branches = 2 if "cfg" in args.mode else 1
        if branches == 2 and not args.cfg:
            raise SystemExit("--mode cfgpair needs --cfg")
        layout = qie_b200.make_layout(world, rank, branches)
        if layout.sp_size > 1:
            runner = qie_b200.UlyssesTransformer(model, layout.sp_group, fused=args.fused)
        g = torch.Generator(device=dev).manual_seed(1)      # one frame: identical inputs on every rank
        lat = torch.randn(NB, N_NOISE, 64, generator=g, device=dev).bfloat16()
        img_lat = torch.randn(NB, N_IMG_TOK - N_NOISE, 64, generator=g, device=dev).bfloat16()
        cond = (torch.randn(NB, T_TXT, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16()
        unc = (torch.randn(NB, T_TXT, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16() if args.cfg else None
        host = {k: v.cpu().pin_memory() for k, v in (("lat", lat), ("img", img_lat), ("cond", cond))}
        if unc is not None:
            host["unc"] = unc.cpu().pin_memory()

    if args.cache and args.mode == "dp":
        sig = qie_b200.flowmatch_sigmas(STEPS_PER_IMAGE, N_NOISE)
        model.cache_schedule([float(qie_b200.model_timestep(float(s_), 1, "cpu")[0]) for s_ in sig[:STEPS_PER_IMAGE]])
        model.cache_prompt("cond", cond)
        if unc is not None:
            model.cache_prompt("uncond", unc)

    shapes_b = IMG_SHAPES * args.batch if args.batch > 1 else IMG_SHAPES

    def denoise(l, i, c, u):
        if args.cache and args.mode == "dp":
            return qie_b200.run_denoise(runner, l, i, c, IMG_SHAPES, STEPS_PER_IMAGE, u, 4.0, use_caches=True)
        if layout is not None and layout.cfg_branches == 2:
            return qie_b200.run_denoise_parallel(runner, layout, l, i, c, u, shapes_b, STEPS_PER_IMAGE, 4.0)
        return qie_b200.run_denoise(runner, l, i, c, shapes_b, STEPS_PER_IMAGE, u, 4.0, batched_cfg=args.batched_cfg)

    def step_resident():
        return denoise(lat, img_lat, cond, unc)

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        res = denoise(d["lat"], d["img"], d["cond"], d.get("unc"))
        out_host.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the edited latents
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    L = qie_b200.lib()
    with ClockSampler(local) as clk:
        n0 = L.qie_launch_count()
        ms_total = timed(step_resident, args.steps)
        launches = L.qie_launch_count() - n0
    clocks = clk.summary()
    if launches == 0 and hasattr(runner, "graphs"):      # sequence-parallel forwards replay CUDA graphs: count one eager step
        runner.graphs, n0 = False, L.qie_launch_count()
        saved = {k: g_.graph for k, g_ in runner._geo.items()}
        for g_ in runner._geo.values():
            g_.graph = None
        step_resident()
        launches = (L.qie_launch_count() - n0) * args.steps
        runner.graphs = True
        for k, g_ in runner._geo.items():
            g_.graph = saved[k]
    ms_step = ms_total / args.steps
    frames = (world if args.mode == "dp" else 1) * args.batch     # frames finished per step by the whole job
    value = frames * 1e3 / ms_step

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    # live per-kernel-class CUDA-event timing of the same step (events on the launching stream, inside qie_forward)
    model.profile(True)      # sequence-parallel modes: rank 0's shard (per-rank work / per-rank time)
    step_resident()
    torch.cuda.synchronize()
    model.read_profile()
    for _ in range(args.steps):
        step_resident()
    torch.cuda.synchronize()
    prof = model.read_profile()
    model.profile(False)
    sus, burst, hbm, which = peaks()
    # dram__bytes_read.sum + dram__bytes_write.sum per gemm_kernel launch: from the ncu --set full capture of THIS round's build
    # (profiles/r02_gemm_traffic.json names the commit it was taken from); null when no capture of the current kernels exists
    traffic, traffic_src = None, None
    tp = ROOT / "profiles" / "r02_gemm_traffic.json"
    if tp.exists() and args.precision == "bf16":
        tj = json.loads(tp.read_text())
        traffic, traffic_src = tj.get("gemm_kernel_avg_dram_bytes_per_launch"), tj.get("source")
    gm, at = prof["gemm"], prof["attention"]
    gemm_tf = gm["work"] / (gm["ms"] * 1e-3) / 1e12 if gm["ms"] else 0.0
    attn_tf = at["work"] / (at["ms"] * 1e-3) / 1e12 if at["ms"] else 0.0
    tot_ms = sum(v["ms"] for v in prof.values())
    # W8A8 modes: the GEMMs run on the 8-bit tensor path, whose peak is measured here (cuBLASLt on the same box, same run);
    # attention stays bf16 and keeps the bf16 peak
    gemm_peak, gemm_burst, gemm_peak_src = sus, burst, f"{which} bf16_tflops_sustained (kernel timed inside a long step); burst {burst}"
    if args.precision != "bf16":
        q8 = q8_library_peak(args.precision, dev)
        if q8 is not None:
            gemm_peak, gemm_burst = q8[0], q8[1]
            gemm_peak_src = f"measured in this run: {q8[2]}, back to back for 2 s (sustained); burst {q8[1]:.1f}"
        else:
            gemm_peak, gemm_burst = 2 * sus, 2 * burst
            gemm_peak_src = f"2 x the {which} bf16 sustained peak (no 8-bit library GEMM in this torch build)"
    roofline = {"bound": "tensor", "kernel": f"gemm_kernel (tcgen05 {args.precision}, all linears of the step)",
                "achieved": gemm_tf, "peak": gemm_peak, "unit": "TFLOP/s", "frac": gemm_tf / gemm_peak, "traffic": traffic,
                "traffic_note": "bytes per launch averaged over the 4 per-block GEMM shapes, ncu --set full "
                                f"({traffic_src or 'no capture of this build'}); algorithmic operand+output bytes average 420 MB per launch",
                "peak_source": gemm_peak_src,
                "launches_per_step": gm["launches"] / args.steps, "avg_launch_ms": gm["ms"] / max(gm["launches"], 1),
                "share_of_step": gm["ms"] / tot_ms if tot_ms else None,
                # the other kernel classes at the top level too (same live event timing; attention against the same tensor peak,
                # adaLN / modulation GEMV against the measured HBM copy peak)
                "attention_achieved": attn_tf, "attention_frac": attn_tf / sus, "attention_frac_of_burst": attn_tf / burst,
                "attention_share_of_step": at["ms"] / tot_ms if tot_ms else None,
                "adaln_achieved_gbs": prof["adaln"]["work"] / (prof["adaln"]["ms"] * 1e-3) / 1e9 if prof["adaln"]["ms"] else 0,
                "adaln_frac": (prof["adaln"]["work"] / (prof["adaln"]["ms"] * 1e-3) / 1e9 / hbm) if prof["adaln"]["ms"] else 0,
                "adaln_share_of_step": prof["adaln"]["ms"] / tot_ms if tot_ms else None,
                "mod_gemv_frac": (prof["mod_gemv"]["work"] / (prof["mod_gemv"]["ms"] * 1e-3) / 1e9 / hbm) if prof["mod_gemv"]["ms"] else 0,
                "frac_of_burst": gemm_tf / gemm_burst,
                "attention": {"achieved": attn_tf, "frac": attn_tf / sus, "share_of_step": at["ms"] / tot_ms if tot_ms else None,
                              "avg_launch_ms": at["ms"] / max(at["launches"], 1)},
                "adaln": {"achieved_gbs": prof["adaln"]["work"] / (prof["adaln"]["ms"] * 1e-3) / 1e9 if prof["adaln"]["ms"] else 0,
                          "peak_gbs": hbm, "share_of_step": prof["adaln"]["ms"] / tot_ms if tot_ms else None},
                "mod_gemv": {"achieved_gbs": prof["mod_gemv"]["work"] / (prof["mod_gemv"]["ms"] * 1e-3) / 1e9 if prof["mod_gemv"]["ms"] else 0,
                             "peak_gbs": hbm, "share_of_step": prof["mod_gemv"]["ms"] / tot_ms if tot_ms else None},
                "event_sum_ms_per_step": tot_ms / args.steps,      # sum of the per-kernel event intervals; the rest of ms_per_step is inter-kernel gaps
                "step_tflops": args.batch * STEPS_PER_IMAGE * (2 if args.cfg else 1) * flops_per_forward(args.layers) / (ms_step * 1e-3) / 1e12}

    strong = None
    if (args.mode == "dp" and not args.no_strong and args.precision == "bf16" and args.workload == "1024x1ref" and not args.cache
            and args.batch == 1 and not args.cfg and args.sched_steps == 0):
        strong = strong_leg(args, model, dev, world, rank, timed)

    eager = None
    if (rank == 0 and world == 1 and not args.no_eager_baseline and args.workload == "1024x1ref" and args.batch == 1
            and args.precision == "bf16" and args.layers == 60):
        e = run_eager_gpu(args, emit=False, min_ms=2000.0)
        eager = {"ms_per_forward": e["dit_forward_ms"], "value": e["value"], "unit": UNIT, "sample": e["sample"],
                 "clocks": e["clocks"], "speedup_of_this_repo": e["ms_per_step"] / ms_step}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        run_sample, scale, desc = cpu_reference_step_factory()
        run_sample()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            run_sample()
        sec_img = (time.perf_counter() - t0) / n * scale
        cpu = {"value": 1.0 / sec_img, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if args.mode == "dp" else "strong", "vs_baseline": None,
                "dtype": {"bf16": "bf16", "fp8": "fp8_e4m3(w8a8)+bf16", "int8": "int8(w8a8)+bf16"}[args.precision], "data": "synthetic",
                "config": workload_config(args, world), "clocks": clocks,
                "e2e": {"value": frames * 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "dit_forward_ms": ms_step / (STEPS_PER_IMAGE * (2 if args.cfg else 1)), "latency_ms_per_batch": ms_step,
                "strong": strong, "gpu_eager_baseline": eager}
        print(json.dumps(line), flush=True)
    failed = bool(strong and strong.get("failed"))
    if world > 1:
        dist.destroy_process_group()
    if failed:
        raise SystemExit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"],
                    help="ours = libqie.so; reference = the reference's CPU path (fp32 oracle on the host cores); eager = the "
                         "oracle in bf16 PyTorch eager on the GPU (what diffusers dispatches), a reported side baseline")
    ap.add_argument("--cfg", action="store_true", help="true-CFG (cond + uncond forwards per step)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp8", "int8"])
    ap.add_argument("--layers", type=int, default=60)
    ap.add_argument("--attn-variant", type=lambda x: int(x, 0), default=0, help="attention kernel variant (0 = library default)")
    ap.add_argument("--cache", action="store_true", help="use the exact schedule/prompt caches (N1); reported separately, "
                    "never the default: the headline recomputes every timestep-dependent vector inside the timed region")
    ap.add_argument("--mode", default="dp", choices=["dp", "cfgpair", "ulysses", "cfg+ulysses"])
    ap.add_argument("--fused", action="store_true", help="ulysses modes: exchange q|k|v and the attention output through "
                    "epilogue stores into peer memory (NVLink) instead of NCCL all-to-alls")
    ap.add_argument("--batch", type=int, default=1, help="frames per forward on every rank (dp) / inside the group (other modes); "
                    "BASELINE configs[4] streams batches of 8")
    ap.add_argument("--batched-cfg", action="store_true", help="with --cfg on one GPU: cond and uncond forward of a step as ONE forward "
                    "of batch 2 (per-element text lengths), the reference's batched_cfg_pipeline.py (README.md:126)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (ONE true-CFG frame over all ranks)")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the bf16 PyTorch-eager GPU side baseline (N = 1)")
    ap.add_argument("--workload", default="1024x1ref", choices=["1024x1ref", "512x2ref"],
                    help="1024x1ref = the headline (BASELINE configs[1]); 512x2ref = one frame of configs[4]: 512x512, two "
                         "reference images (3072 image tokens), 448 text tokens, 4 steps (use with --cfg)")
    ap.add_argument("--sched-steps", type=int, default=0, help="denoise steps per image (0 = the workload's own: 2 for the "
                    "headline; BASELINE configs[2] / configs[3] use 4)")
    args = ap.parse_args()
    global METRIC, IMG_SHAPES, N_NOISE, N_IMG_TOK, T_TXT, STEPS_PER_IMAGE
    if args.workload == "512x2ref":
        METRIC = "edited_512x512_two_image_frames_per_s_4step"
        IMG_SHAPES = [[(1, 32, 32), (1, 32, 32), (1, 32, 32)]]
        N_NOISE, N_IMG_TOK, T_TXT, STEPS_PER_IMAGE = 1024, 3072, 448, 4
    if args.sched_steps > 0:
        STEPS_PER_IMAGE = args.sched_steps
        if args.workload == "1024x1ref":
            METRIC = f"edited_1024x1024_images_per_s_{args.sched_steps}step"
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager":
        run_eager_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
