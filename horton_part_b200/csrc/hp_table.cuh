// Tabulated (piecewise-cubic) basis functions on the molecular grid: basis_type="numeric" of gLISA
// (NumericBasisFuncHelper, core/basis.py:330-387: every shell of an element is a not-a-knot CubicSpline on the
// element's radial knots, evaluated with extrapolation).  All shells of an atom share the atom's knots, so
// one interval look-up per (atom, point) serves every shell; shell m owns 4 x (n_a - 1) PPoly coefficients
// (highest power first) at shell_coef + shell_coef_off[m].  Interval look-up = the table + forward scan of
// the spline pass (hp_spline.cu): searchsorted_right(x, r) - 1 clamped to [0, n - 2].
#pragma once

#include "hp_common.cuh"

namespace hp {

constexpr int kTabLutShift = 15;  // hi32(r) >> 15: sign, exponent, 5 mantissa bits (kLutShift in hp_spline.cu)

struct TableArgs {
    const int* knot_off;             // natom + 1
    const double* knots;             // concatenated per atom
    const int* lut_meta;             // 3 per atom: key0, nbins, offset into lut
    const unsigned short* lut;       // pooled interval tables
    const long long* shell_coef_off; // per shell: offset of its coefficient block in shell_coef
    const double* shell_coef;
};

// interval of r on atom a's knots and the offset d = r - x[i]
__device__ __forceinline__ int table_interval(const TableArgs& t, int a, double r, double& d) {
    const int o = t.knot_off[a], n = t.knot_off[a + 1] - o;
    const double* xk = t.knots + o;
    int bin = (__double2hiint(r) >> kTabLutShift) - t.lut_meta[3 * a];
    bin = max(0, min(bin, t.lut_meta[3 * a + 1] - 1));
    int i = t.lut[t.lut_meta[3 * a + 2] + bin];
    const int last = n - 2;
    while (i < last && xk[i + 1] <= r) ++i;
    d = r - xk[i];
    return i;
}

// PPoly's evaluation order: c3 + c2 d + c1 d^2 + c0 d^3
__device__ __forceinline__ double table_cubic(const double* __restrict__ c, double d) {
    double z = d;
    double res = fma(c[2], z, c[3]);
    z *= d;
    res = fma(c[1], z, res);
    z *= d;
    return fma(c[0], z, res);
}

}  // namespace hp
