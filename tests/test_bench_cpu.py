"""CPU suite: the measurement contract of bench.py that can be checked without a GPU — the algorithmic FLOP count the
roofline is computed from (SURVEY §8d known answers), the clock-sample parser, and that the CUDA arm refuses to run without a
device instead of falling back to the CPU."""
import importlib.util
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_flops_match_survey_8d(bench):
    # config 2: 167.45 TFLOP per forward (linear 114.80 + attention 52.62); T = 219: 166.49; config 3 uncond (T=200): 165.99
    assert bench.flops_per_forward() / 1e12 == pytest.approx(167.42, abs=0.05)      # without the 0.03 TFLOP "top"
    assert bench.flops_per_forward(T=219) / 1e12 == pytest.approx(166.46, abs=0.05)
    assert bench.flops_per_forward(T=200) / 1e12 == pytest.approx(165.96, abs=0.05)
    # 512^2 with one reference image (S = 2304): 35.24;  config 5 (B = 8, S = 3520): 455.97
    assert bench.flops_per_forward(S_i=2048, T=256) / 1e12 == pytest.approx(35.24, abs=0.05)
    assert bench.flops_per_forward(S_i=3072, T=448, B=8) / 1e12 == pytest.approx(455.97, abs=0.3)
    # per block per token: 226 492 416 linear FLOP
    assert 2 * 3072 * 12 * 3072 == 226_492_416


def test_clock_sampler_summary_parses_nvidia_smi_rows(bench):
    s = bench.ClockSampler(0)
    s.rows = [["1455", "1965", "997.63", "Not Active", "Not Active", "Not Active", "Active"],
              ["1470", "1965", "1001.2", "Not Active", "Not Active", "Not Active", "Active"],
              ["1440", "1965", "990.0", "Not Active", "Not Active", "Not Active", "Not Active"],
              ["garbage"]]
    out = s.summary()
    assert out["sm_mhz"] == 1455.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["power_w_max"] == pytest.approx(1001.2)
    s.rows = []
    assert bench.ClockSampler(0).summary()["reasons"] == ["unavailable"]


def test_cuda_arm_refuses_to_run_without_a_device(bench, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "1"])
    with pytest.raises(SystemExit) as e:
        bench.main()
    assert "no CPU fallback" in str(e.value)


def test_eight_bit_peak_probe_reports_absence_instead_of_guessing(bench):
    """The W8A8 lines hold their GEMMs against a cuBLASLt e4m3 / int8 peak measured in the same run; where that cannot be
    measured (no CUDA device here) the probe returns None and the line says it falls back to 2 x the bf16 peak."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    assert bench.q8_library_peak("fp8", "cuda:0", seconds=0.01) is None
    assert bench.q8_library_peak("int8", "cuda:0", seconds=0.01) is None


def test_strong_record_verdict(bench):
    """bench.strong_check: the sequence-parallel frame is held to 1e-2 (`parity_ok`) and fails the run beyond the north star's
    2e-2, on a barrier time-out, a NaN or a cosine under 0.999; the CFG pair must be bit-identical."""
    def rec(err, cos=0.9999, timeouts=0, final=0.0, tol=1e-2):
        return {"parity_err": err, "parity_tolerance": tol, "barrier_timeouts": timeouts, "final_latent_cosine": cos,
                "final_latent_max_rel_err": final}
    ok = bench.strong_check(rec(7.0e-3, final=6e-3), sp=4)
    assert ok["parity_ok"] and "failed" not in ok and ok["parity_hard_limit"] == 2e-2
    soft = bench.strong_check(rec(1.3e-2, final=9e-3), sp=4)
    assert not soft["parity_ok"] and "failed" not in soft
    for bad in (rec(2.5e-2), rec(float("nan")), rec(5e-3, timeouts=1), rec(5e-3, cos=0.99)):
        assert "failed" in bench.strong_check(bad, sp=2)
    pair = bench.strong_check(rec(0.0, cos=1.0, final=0.0, tol=0.0), sp=1)
    assert pair["parity_ok"] and "failed" not in pair
    assert "failed" in bench.strong_check(rec(1e-6, cos=1.0, final=0.0, tol=0.0), sp=1)       # a CFG pair that is not bit-identical
    assert "failed" in bench.strong_check(rec(0.0, cos=1.0, final=1e-6, tol=0.0), sp=1)


def _run_bench(args, env_extra):
    import os
    import subprocess
    env = dict(os.environ, QIE_BENCH_CPU_SAMPLE_S="0.2", **env_extra)
    return subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: ONE JSON line on rank 0 with the keys the driver reads (same metric / unit / config as the
    CUDA arm, impl = reference, cpu_baseline describing this run, e2e repeating the value with zero copy bytes)."""
    import json
    r = _run_bench(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"], {})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "edited_1024x1024_images_per_s_2step_lightning" and d["unit"] == "img/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] == pytest.approx(1e3 / d["value"], rel=1e-6)
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "full-width" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("Qwen-Image-Edit-2509 MMDiT denoise") and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                   {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_strong_mode_partition_by_world_size():
    """the strong-scaling leg of bench.py: which partition ONE true-CFG frame takes at each GPU count (north star: CFG pair at 2,
    CFG pair x Ulysses at 4 / 8; the sequence-parallel degree must divide the 24 heads)"""
    import bench
    assert bench.strong_mode(1) == ("single", 1, 1)
    assert bench.strong_mode(2) == ("cfgpair", 2, 1)
    assert bench.strong_mode(4)[1:] == (2, 2) and bench.strong_mode(8)[1:] == (2, 4) and "fused-ulysses4" in bench.strong_mode(8)[0]
    assert bench.strong_mode(6)[1:] == (2, 3) and bench.strong_mode(3)[1:] == (1, 3)
    assert bench.strong_mode(10)[0] is None and bench.strong_mode(7)[0] is None      # 5 and 7 do not divide 24 heads
