// Pieces shared by the dense and the cut-off promolecule kernels.
#pragma once

#include "hp_common.cuh"
#include "hp_math.cuh"

namespace hp {

constexpr int kTileAtoms = 128;         // atoms per shared-memory tile
constexpr int kTileShells = 1024;       // shells per shared-memory tile
constexpr int kMaxPartials = 4096;      // size of the entropy partial-sum buffer

struct __align__(16) AtomRec {
    double x, y, z;
    int s0, ns;  // first shell (tile-relative) and shell count
};

// Pro-atom density of one atom at kP points (squared distances d2[]).  Shell parameters are read
// once per shell for all points.
template <int F, bool FAST = false>
__device__ __forceinline__ double shell_value(double2 ab, double n, double r) {
    if (F == HP_FUNCTOR_GENERAL) {
        const double rn = (n == 1.0) ? r : ((n == 2.0) ? r * r : pow(r, n));
        return exp(-ab.y * rn);
    }
    return exp_neg_poly<!FAST>(-ab.y * r);
}

// Pro-atom density of one atom at kP points (squared distances d2[]).  Shell parameters are read
// once per shell for all points; the first shell's (A, alpha) arrives in registers (ab0).
// FAST: the caller guarantees non-zero distances and exponent arguments above -700 for every point
// (chunk geometry), so the zero-distance and underflow guards are dropped and the 5-op sqrt is used.
template <int F, int kP, bool FAST = false>
__device__ __forceinline__ void eval_proatom(const double (&d2)[kP], int s0, int ns,
                                             const double2* __restrict__ sAB,
                                             const double* __restrict__ sN, double (&f)[kP],
                                             double2 ab0) {
    double r[kP];
#pragma unroll
    for (int j = 0; j < kP; ++j) r[j] = (F == HP_FUNCTOR_GAUSS) ? d2[j] : (FAST ? sqrt_fast(d2[j]) : sqrt_nocall(d2[j]));
    if (ns > 0) {
        const double n0 = (F == HP_FUNCTOR_GENERAL) ? sN[s0] : 1.0;
#pragma unroll
        for (int j = 0; j < kP; ++j) f[j] = ab0.x * shell_value<F, FAST>(ab0, n0, r[j]);
    } else {
#pragma unroll
        for (int j = 0; j < kP; ++j) f[j] = 0.0;
    }
    for (int k = 1; k < ns; ++k) {
        const double2 ab = sAB[s0 + k];  // (A, alpha)
        const double n = (F == HP_FUNCTOR_GENERAL) ? sN[s0 + k] : 1.0;
#pragma unroll
        for (int j = 0; j < kP; ++j) f[j] = fma(ab.x, shell_value<F, FAST>(ab, n, r[j]), f[j]);
    }
}

}  // namespace hp
