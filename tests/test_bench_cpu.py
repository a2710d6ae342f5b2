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
