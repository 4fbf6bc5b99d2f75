#!/usr/bin/env python
"""Coefficients and accuracy estimate of the base-2 exponential variant of the hot kernel
(HP_LOC_EXP2, off by default).

    exp(-alpha r) = 2^y,   y = -(alpha log2 e) r,   y = k + s,  k = round(y),  |s| <= 1/2
    2^s = 1 + s (ln2 + s G(s)),   G(s) = ln2^2 g(ln2 s),   G_i = g_i ln2^(i+2)

with g the degree-9 polynomial of hp_math.cuh (exp(t) = 1 + t + t^2 g(t) on |t| <= ln2/2).  The
reduction needs no Cody-Waite constants (s = y - k is exact), which saves one FP64 operation per
shell evaluation; the price is a second rounding in the argument (beta = fl(alpha log2 e)).

Prints the G_i as C literals and the error of the scheme in ulp against a 50-digit reference on
random arguments, with the FMAs emulated in extended precision (no GPU needed).
"""

from decimal import Decimal, getcontext

import numpy as np

getcontext().prec = 60
G_EXP = ["0.5000000000000001", "0.16666666666666669", "0.04166666666662413", "0.008333333333330062",
         "0.0013888888917213717", "0.00019841269863053618", "2.4801521295954376e-05", "2.7557268459997064e-06",
         "2.7620088445409746e-07", "2.510038549551032e-08"]  # HP_EXPG0..9 (hp_math.cuh)
LN2 = Decimal(2).ln()


def coefficients():
    # the literals in hp_math.cuh ARE doubles: take their exact binary values
    g = [Decimal(float(v)) for v in G_EXP]
    return [float(gi * LN2 ** (i + 2)) for i, gi in enumerate(g)]


def fma(a, b, c):
    return np.float64(np.longdouble(a) * np.longdouble(b) + np.longdouble(c))


def exp2_scheme(y, G):
    magic = np.float64(6755399441055744.0)
    t = y + magic
    kd = t - magic
    s = y - kd
    g = fma(G[9], s, G[8])
    for i in range(7, -1, -1):
        g = fma(g, s, G[i])
    p = fma(g, s, np.float64(float(LN2)))
    p = fma(p, s, np.float64(1.0))
    return np.ldexp(p, kd.astype(np.int64))


def main():
    G = coefficients()
    print("// G_i = g_i * ln2^(i+2), correctly rounded")
    for i, v in enumerate(G):
        print(f"#define HP_EXP2G{i} {v!r}")
    print(f"// ln2   = {float(LN2)!r}\n// log2e = {float(1 / LN2)!r}")
    rng = np.random.default_rng(0)
    y = -np.concatenate([rng.uniform(0, 2, 4000), rng.uniform(0, 60, 4000), rng.uniform(0, 1000, 2000)])
    got = exp2_scheme(y.astype(np.float64), [np.float64(v) for v in G])
    worst = 0.0
    for yi, gi in zip(y, got):
        ref = (Decimal(float(yi)) * LN2).exp()
        ulp = Decimal(float(np.spacing(gi)))
        worst = max(worst, float(abs(Decimal(float(gi)) - ref) / ulp))
    print(f"// max error of 2^y over {len(y)} arguments in (-1000, 0]: {worst:.3f} ulp (argument taken as exact)")


if __name__ == "__main__":
    main()
