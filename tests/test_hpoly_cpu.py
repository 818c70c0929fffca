"""The attention kernel's packed-half polynomial exponential (csrc/attn_tc.cu: ex2_hpoly), restated operation by
operation with fp16 roundings in tools/fit_hpoly.py: accuracy bounds quoted in the kernel's comments and the edge cases
of the range reduction.  (The CUDA path itself is checked against the oracle in tests/test_kernels_gpu.py.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fit_hpoly as hp  # noqa: E402


def test_polynomial_accuracy_over_every_fp16_fraction():
    mx, rms, mean = hp.fraction_error()
    mx_r, rms_r = hp.rounding_error()
    assert mx < 5.7e-4 and rms < 2.1e-4 and abs(mean) < 1e-6, (mx, rms, mean)
    # within 15 % of what a correctly rounded fp16 exponential achieves on the same points
    assert mx < 1.15 * mx_r and rms < 1.02 * rms_r, (mx, mx_r, rms, rms_r)


def test_whole_sequence_over_the_softmax_range():
    # arguments of the softmax: (-inf, ~8] (the lazy-rescale threshold); normal fp16 results down to 2^-14
    mx, rms = hp.end_to_end_error(-14.0, 9.0)
    assert mx < 6e-4 and rms < 2.2e-4, (mx, rms)


def test_range_reduction_edges():
    x = np.array([0.0, -0.5, 0.5, -1.0, 1.0, 8.0, 9.49, -13.51, -14.49, -15.0, -16.0, -100.0, -1e30, -np.inf], dtype=np.float32)
    y = hp.ex2_hpoly(x).astype(np.float64)
    assert y[0] == 1.0 and y[3] == 0.5 and y[4] == 2.0 and y[5] == 256.0
    assert abs(y[1] / 2 ** -0.5 - 1) < 6e-4 and abs(y[2] / 2 ** 0.5 - 1) < 6e-4
    assert abs(y[6] / 2 ** float(np.float16(9.49)) - 1) < 6e-4          # n = 9: exponent field 24, still finite
    assert abs(y[7] / 2 ** float(np.float16(-13.51)) - 1) < 2e-3         # sub-normal result (n = -14, r < 1)
    assert np.all(y[9:] == 0.0)                                          # clamp at -15: biased exponent 0 -> exactly 0
    assert np.all(np.isfinite(y)) and np.all(y >= 0)
    # monotone over the whole range
    xs = np.linspace(-14, 9, 20001).astype(np.float32)
    ys = hp.ex2_hpoly(xs).astype(np.float64)
    assert np.all(np.diff(ys) >= -1e-3 * ys[1:])   # monotone up to one fp16 ulp at the interval joins


def test_constants_match_the_kernel_source():
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaussctrl_b200", "csrc",
                            "attn_tc.cu")).read()
    for bits in hp.COEF_BITS:
        assert ("0x%04X%04Xu" % (bits, bits)) in src, hex(bits)
    assert "0x660F660Fu" in src and "0xCB80CB80u" in src and "0x7C007C00" in src   # 1551, -15, exponent mask
