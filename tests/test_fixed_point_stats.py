"""Host-side model of the BatchNorm statistic accumulators of the tcgen05 layer kernel
(papc_b200/csrc/sa_mlp_tt.cu, TtArgs::fix_acc): every CTA adds its fp64 partial sum p as the pair
(I, F) = (trunc(p), rint((p - trunc(p)) * 2^54)) with 64-bit INTEGER atomics and the last CTA
reconstructs (double)sum(I) + (double)sum(F) * 2^-54.  The properties the kernel relies on:
the split is exact to 2^-54 absolute, integer addition is associative (any arrival order of the
CTAs gives the same words, hence run-to-run determinism), and 296 CTAs cannot overflow the words
as long as |p| < 2^53 (larger or NaN partials raise the fallback flag instead)."""
import math

import numpy as np
import pytest

MASK = (1 << 64) - 1


def split(p):
    ip = math.trunc(p)
    return ip, int(np.rint((p - ip) * 2.0 ** 54))


def accumulate(parts, order):
    acc_i = acc_f = 0
    for k in order:                      # unsigned 64-bit wrap-around adds, as atomicAdd(unsigned long long)
        i, f = split(float(parts[k]))
        acc_i = (acc_i + i) & MASK
        acc_f = (acc_f + f) & MASK
    to_signed = lambda v: v - (1 << 64) if v >> 63 else v
    return to_signed(acc_i), to_signed(acc_f)


@pytest.mark.parametrize("scale", [1e-6, 1.0, 1e4, 1e12])
def test_any_arrival_order_gives_the_same_words_and_the_exact_sum(scale):
    rng = np.random.default_rng(int(math.log10(scale)) + 20)
    parts = rng.standard_normal(296) * scale            # grid_rows cap: 2 x 148 CTAs
    ref_words = accumulate(parts, range(len(parts)))
    for _ in range(5):
        assert accumulate(parts, rng.permutation(len(parts))) == ref_words
    total = float(ref_words[0]) + float(ref_words[1]) * 2.0 ** -54
    exact = math.fsum(parts.tolist())
    # every partial is truncated to a multiple of 2^-54 (<= 2^-55 each), then one rounding to double
    assert abs(total - exact) <= len(parts) * 2.0 ** -55 + abs(exact) * 2.0 ** -52


def test_words_cannot_overflow_below_the_flag_threshold():
    p = math.nextafter(2.0 ** 53, 0.0)                  # largest partial the kernel accepts
    i, f = split(p)
    assert f == 0 and 296 * abs(i) < 2 ** 63            # integer words
    i, f = split(0.9999999999999999)
    assert i == 0 and 296 * abs(f) < 2 ** 63            # fraction words (|F| <= 2^54)
    i, f = split(-123.75)
    assert (i, f) == (-123, -(3 << 52))                 # trunc toward zero, fraction carries the sign
