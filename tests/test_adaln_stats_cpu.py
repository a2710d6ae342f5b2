"""CPU suite: the row statistics of the CTA-row adaLN kernel (csrc/glue.cu, ln_mod_cta_kernel), restated in numpy fp32.

The eight warps of a CTA each reduce their D/8-column slice of a row to (mean_w, M2_w = sum (x - mean_w)^2); the eight pairs are
merged with the pairwise formula of Chan, Golub & LeVeque:  mean = avg(mean_w),  M2 = sum M2_w + (D/8) * sum (mean_w - mean)^2.
This file checks, without a GPU, that the merge is the two-pass LayerNorm variance up to fp32 rounding — also for rows whose mean
is large against their spread, where the one-pass E[x^2] - mean^2 form loses every digit — i.e. that the kernel may replace the
warp-per-row two-pass kernels inside the bf16 output tolerance (2^-8) the GPU tests hold all of them to."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

F = np.float32
WARPS = 8


def merged_stats(x: np.ndarray, eps: float = 1e-6):
    """fp32 restatement of the kernel's statistics for one row x[D] (D a multiple of 8)."""
    D = x.shape[0]
    w = D // WARPS
    sl = x.astype(F).reshape(WARPS, w)
    mean_w = (sl.sum(axis=1, dtype=F) * F(1.0 / w)).astype(F)
    m2_w = ((sl - mean_w[:, None]) ** 2).sum(axis=1, dtype=F)
    mean = (mean_w.sum(dtype=F) * F(0.125)).astype(F)
    dev = ((mean_w - mean) ** 2).sum(dtype=F)
    var = (m2_w.sum(dtype=F) + F(w) * dev) * F(1.0 / D)
    return mean, F(1.0) / np.sqrt(var + F(eps), dtype=F)


def two_pass_f64(x: np.ndarray, eps: float = 1e-6):
    x = x.astype(np.float64)
    mean = x.mean()
    return mean, 1.0 / np.sqrt(((x - mean) ** 2).mean() + eps)


@pytest.mark.parametrize("D", [1024, 2048, 3072])
@pytest.mark.parametrize("offset,spread", [(0.0, 1.0), (0.5, 3.0), (100.0, 1.0), (-3000.0, 0.5), (0.0, 1e-3)])
def test_merged_statistics_equal_two_pass_variance(D, offset, spread):
    rng = np.random.default_rng(D + int(abs(offset)))
    x = (offset + spread * rng.standard_normal(D)).astype(F)
    x[7] += 30 * spread                                        # an outlier channel, like the massive activations of the text stream
    mean, rstd = merged_stats(x)
    m64, r64 = two_pass_f64(x)
    y, y64 = (x - mean) * rstd, (x.astype(np.float64) - m64) * r64
    # far inside the bf16 rounding of the stored output (2^-8 relative to the row maximum)
    assert np.abs(y - y64).max() <= 2.0 ** -8 * np.abs(y64).max() * 1e-2
    var64 = ((x.astype(np.float64) - m64) ** 2).mean()
    merged_err = abs((1.0 / float(rstd) ** 2 - 1e-6) - var64)
    assert merged_err <= 2e-5 * var64
    if abs(offset) >= 100:                                      # the one-pass form the merge avoids loses its digits here
        xf = x.astype(F)
        one_pass = (xf * xf).mean(dtype=F) - xf.mean(dtype=F) ** 2
        assert abs(float(one_pass) - var64) > 20 * merged_err


@settings(max_examples=100, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.floats(-50, 50), st.floats(0.01, 20), st.integers(0, 7), st.floats(0, 40))
def test_merged_statistics_property(seed, offset, spread, hot_slice, slice_shift):
    """slices with different means (one warp's columns shifted): the between-slice term of the merge carries the variance"""
    D = 3072
    rng = np.random.default_rng(seed)
    x = (offset + spread * rng.standard_normal(D)).astype(F)
    x.reshape(WARPS, -1)[hot_slice] += F(slice_shift)
    mean, rstd = merged_stats(x)
    m64, r64 = two_pass_f64(x)
    assert abs(float(mean) - m64) <= 1e-5 * (abs(m64) + spread + slice_shift)
    # fp32 slice means carry ~1e-7 relative error of the OFFSET; where the spread is 1e-3 of the offset that is a few 1e-5 of rstd
    # (first-order in the between-slice term) — still 20x under half a bf16 ulp (2^-9) of the stored output
    assert abs(float(rstd) - r64) <= 1e-4 * r64
