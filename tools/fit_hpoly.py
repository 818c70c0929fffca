"""The packed-half polynomial exponential of the attention kernel (csrc/attn_tc.cu: ex2_hpoly), restated in numpy with
the kernel's exact operation sequence and fp16 roundings, and the search that picked its coefficients.

  python tools/fit_hpoly.py          accuracy of the shipped coefficients (every fp16 fraction, and x in [-16, 9])
  python tools/fit_hpoly.py search   re-run the neighbourhood search around the fp32 minimax coefficients
"""
import sys

import numpy as np

F16 = np.float16
# fp16 bit patterns used by the kernel: c3, c2, c1, c0 of 2^f ~ ((c3 f + c2) f + c1) f + c0 on [-0.5, 0.5]
COEF_BITS = (0x2B08, 0x33C0, 0x398C, 0x3C00)
MAGIC = 1551.0      # 1536 + 15: fp16 ulp is 1 in [1024, 2048), and the low 5 bits of the sum's pattern = n + 15
CLAMP = -15.0


def bits_to_f16(b):
    return np.array([b], dtype=np.uint16).view(F16)[0]


COEF = tuple(float(bits_to_f16(b)) for b in COEF_BITS)


def fma16(a, b, c):
    """fma.rn.f16: one rounding"""
    return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(F16) if np.isscalar(b) else \
        (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(F16)


def horner(f, coef=COEF):
    c3, c2, c1, c0 = coef
    r = (f.astype(np.float64) * c3 + c2).astype(F16)
    r = (r.astype(np.float64) * f.astype(np.float64) + c1).astype(F16)
    r = (r.astype(np.float64) * f.astype(np.float64) + c0).astype(F16)
    return r


def ex2_hpoly(x32, coef=COEF):
    """x: float32 array -> fp16 array, the kernel's sequence: cvt.rn.f16, max(-15), +1551, -1551, h - n, (fi << 10) &
    0x7C00 as the bit pattern of 2^n, three fma.rn.f16, one mul.rn.f16."""
    h = np.maximum(x32.astype(F16), F16(CLAMP))
    fi = (h.astype(np.float32) + np.float32(MAGIC)).astype(F16)          # exact integer 1551 + round(h)
    n = (fi.astype(np.float32) - np.float32(MAGIC)).astype(F16)
    f = (h.astype(np.float32) - n.astype(np.float32)).astype(F16)        # exact, in [-0.5, 0.5]
    scale = ((fi.view(np.uint16).astype(np.uint32) << 10) & 0x7C00).astype(np.uint16).view(F16)
    r = horner(f, coef)
    return (r.astype(np.float64) * scale.astype(np.float64)).astype(F16)


def all_fractions():
    allh = np.arange(0, 65536, dtype=np.uint32).astype(np.uint16).view(F16)
    return allh[np.isfinite(allh) & (np.abs(allh.astype(np.float32)) <= 0.5)]


def fraction_error(coef=COEF):
    f = all_fractions()
    true = np.exp2(f.astype(np.float64))
    rel = (horner(f, coef).astype(np.float64) - true) / true
    return float(np.abs(rel).max()), float(np.sqrt((rel ** 2).mean())), float(rel.mean())


def rounding_error():
    f = all_fractions()
    true = np.exp2(f.astype(np.float64))
    rel = (true.astype(F16).astype(np.float64) - true) / true
    return float(np.abs(rel).max()), float(np.sqrt((rel ** 2).mean()))


def end_to_end_error(lo=-14.0, hi=9.0, n=400001):
    """against 2^x of the fp16-rounded argument (the argument rounding is the caller's, common to any fp16 path)"""
    x = np.linspace(lo, hi, n).astype(np.float32)
    got = ex2_hpoly(x).astype(np.float64)
    want = np.exp2(x.astype(F16).astype(np.float64))
    rel = (got - want) / want
    return float(np.abs(rel).max()), float(np.sqrt((rel ** 2).mean()))


def search():
    base = [0.0551716648, 0.2426111251, 0.6932609677, 0.9999280572]   # fp32 degree-3 minimax of 2^f on [-0.5, 0.5]

    def nb(v, k):
        u = int(np.array([v], dtype=F16).view(np.uint16)[0])
        return [float(np.array([u + d], dtype=np.uint16).view(F16)[0]) for d in range(-k, k + 1)]
    best = (9.0, None)
    for c3 in nb(base[0], 8):
        for c2 in nb(base[1], 4):
            for c1 in nb(base[2], 3):
                for c0 in nb(base[3], 1):
                    e = fraction_error((c3, c2, c1, c0))
                    if e[0] < best[0]:
                        best = (e[0], (c3, c2, c1, c0), e)
    print("best max rel %.3e rms %.3e mean %.1e" % best[2])
    for v in best[1]:
        print("  %.10f  0x%04X" % (v, int(np.array([v], dtype=F16).view(np.uint16)[0])))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "search":
        search()
    else:
        print("coefficients:", ", ".join("%.10f (0x%04X)" % (c, b) for c, b in zip(COEF, COEF_BITS)))
        print("fractions in [-0.5, 0.5]: max rel %.3e, rms %.3e, mean %.1e" % fraction_error())
        print("correctly rounded fp16 exp2 on the same points: max rel %.3e, rms %.3e" % rounding_error())
        print("x in [-14, 9] (whole sequence, vs 2^fp16(x)): max rel %.3e, rms %.3e" % end_to_end_error())
