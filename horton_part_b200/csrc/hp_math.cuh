// FP64 device math for the hot kernels: no-call sqrt and exp for non-positive arguments.
// Accuracy of every routine is measured by tools/fp64_probe.cu (run on the B200; results quoted
// in DESIGN.md).
#pragma once

#include <cuda_runtime.h>

namespace hp {

// exp(r) = 1 + r + r^2 g(r) on |r| <= ln2/2; g interpolated at Chebyshev nodes (degree 9),
// max relative error of the polynomial 1.6e-17.
// Literals (not a __constant__ array): ptxas then feeds them to DFMA straight from the immediate
// constant bank instead of re-loading them into registers in every iteration of the atom loop.
#define HP_EXPG0 0.5000000000000001
#define HP_EXPG1 0.16666666666666669
#define HP_EXPG2 0.04166666666662413
#define HP_EXPG3 0.008333333333330062
#define HP_EXPG4 0.0013888888917213717
#define HP_EXPG5 0.00019841269863053618
#define HP_EXPG6 2.4801521295954376e-05
#define HP_EXPG7 2.7557268459997064e-06
#define HP_EXPG8 2.7620088445409746e-07
#define HP_EXPG9 2.510038549551032e-08

__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// sqrt of a finite d2 >= 0 without a slow-path call: MUFU.RSQ64H seed (2^-20 measured), one
// coupled Goldschmidt step, one Heron step = 1 DMUL + 5 DFMA.  d2 below the normal range (a grid
// point sitting on a nucleus) -> 0.
__device__ __forceinline__ double sqrt_nocall(double d2) {
    const double y = rsqrt_seed(d2);
    double g = d2 * y;                                                               // ~sqrt(d2)
    double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));  // y/2
    const double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    const double r = fma(fma(-g, g, d2), h, g);  // Heron step
    return (__double2hiint(d2) < 0x00100000) ? 0.0 : r;
}

// Same seed, two Heron steps with the seed's y/2 as the (approximate) reciprocal: errors 2^-20 ->
// 2^-40 -> 2^-60 before the final rounding = 1 DMUL + 4 DFMA.  Only for d2 in the normal range
// (callers guarantee a non-zero distance): no zero/denormal guard.
__device__ __forceinline__ double sqrt_fast(double d2) {
    const double y = rsqrt_seed(d2);
    const double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));  // y/2
    double g = d2 * y;
    g = fma(fma(-g, g, d2), h, g);
    return fma(fma(-g, g, d2), h, g);
}

// Polynomial / reduction constants of the atom loop, pinned in registers: they are read once per
// thread through an opaque (asm volatile) global load, which makes them run-time values to ptxas.
// Spelled as literals each one is re-materialised with two moves at every use, and read from a
// __constant__ table they are re-fetched (13 LDC.64) in every iteration of the atom loop.
// c_exp_regs keeps a __constant__ copy for the probes in tools/.
#define HP_EXP_TABLE                                                                                   \
    {HP_EXPG0, HP_EXPG1, HP_EXPG2, HP_EXPG3, HP_EXPG4, HP_EXPG5, HP_EXPG6, HP_EXPG7, HP_EXPG8, HP_EXPG9, \
     1.4426950408889634, -6.93147180559945286e-01, -2.31904681384629956e-17}
__constant__ double c_exp_regs[13] = HP_EXP_TABLE;
__device__ double d_exp_regs[13] = HP_EXP_TABLE;

struct ExpConsts {
    double g[10], log2e, ln2hi, ln2lo;
    __device__ __forceinline__ void load() {
        double v[13];
        const unsigned long long base = static_cast<unsigned long long>(__cvta_generic_to_global(d_exp_regs));
#pragma unroll
        for (int i = 0; i < 13; ++i) asm volatile("ld.global.f64 %0, [%1];" : "=d"(v[i]) : "l"(base + 8ull * i));
#pragma unroll
        for (int i = 0; i < 10; ++i) g[i] = v[i];
        log2e = v[10];
        ln2hi = v[11];
        ln2lo = v[12];
    }
};

// Unguarded exp(x), -700 < x <= 0, constants from registers (same arithmetic as exp_neg_poly).
// TWO_STEP = false drops the second Cody-Waite constant: r = x - k fl(ln2) in one FMA (the product is
// exact inside the FMA), leaving the systematic error k (ln2 - fl(ln2)) = k * 2.3e-17 in the reduced
// argument, i.e. a RELATIVE error of the result of at most 0.1 ulp per unit of k (1.4e-15 at
// x = -40, 2.4e-14 at the underflow edge) on top of the polynomial's 0.67 ulp -- 15 FP64 operations
// instead of 16, and the argument x = fl(alpha r) still rounds exactly like the reference's
// np.exp(-S * r) (mbis.py:286), which the base-2 variant below does not.
template <bool TWO_STEP = true>
__device__ __forceinline__ double exp_neg_poly_regs(double x, const ExpConsts& c) {
    const double t = fma(x, c.log2e, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, c.ln2hi, x);
    if (TWO_STEP) r = fma(kd, c.ln2lo, r);
    double g = fma(c.g[9], r, c.g[8]);
#pragma unroll
    for (int i = 7; i >= 0; --i) g = fma(g, r, c.g[i]);
    double p = fma(g, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// ---- base-2 variant (HP_LOC_EXP2, tools/exp2_coeffs.py) -----------------------------------------
// exp(-alpha r) = 2^y with y = -(alpha log2 e) r: the reduction y = k + s is exact without
// Cody-Waite constants, 2^s = 1 + s (ln2 + s G(s)) with G_i = g_i ln2^(i+2).  15 FP64 operations
// instead of 16; <= 0.87 ulp for an exact argument, and the argument carries a second rounding
// (beta = fl(alpha log2 e)).  Off by default: kept for A/B runs on the GPU.
#define HP_EXP2_TABLE                                                                                  \
    {0.24022650695910078, 0.05550410866482158, 0.009618129107618658, 0.0013333558146423209,            \
     0.0001540353042479538, 1.5252733820805852e-05, 1.3215451619146607e-06, 1.0178067259946989e-07,   \
     7.0709810833657675e-09, 4.454105125293416e-10, 0.6931471805599453, 1.4426950408889634}
__device__ double d_exp2_regs[12] = HP_EXP2_TABLE;

struct Exp2Consts {
    double g[10], ln2, log2e;
    __device__ __forceinline__ void load() {
        double v[12];
        const unsigned long long base = static_cast<unsigned long long>(__cvta_generic_to_global(d_exp2_regs));
#pragma unroll
        for (int i = 0; i < 12; ++i) asm volatile("ld.global.f64 %0, [%1];" : "=d"(v[i]) : "l"(base + 8ull * i));
#pragma unroll
        for (int i = 0; i < 10; ++i) g[i] = v[i];
        ln2 = v[10];
        log2e = v[11];
    }
};

// 2^y for -1021 < y <= 0, constants from registers.
__device__ __forceinline__ double exp2_neg_poly_regs(double y, const Exp2Consts& c) {
    const double t = y + 6755399441055744.0;
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    const double s = y - kd;  // exact
    double g = fma(c.g[9], s, c.g[8]);
#pragma unroll
    for (int i = 7; i >= 0; --i) g = fma(g, s, c.g[i]);
    double p = fma(g, s, c.ln2);
    p = fma(p, s, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// y <= -1021 (or a negative NaN): 2^y at or below 4.5e-308 is flushed to exactly 0 (see exp_arg_tiny).
__device__ __forceinline__ bool exp2_arg_tiny(double y) {
    return static_cast<unsigned>(__double2hiint(y)) >= 0xC08FE800u;
}

// x <= -708 (or a negative NaN): results at or below 3.4e-308 are flushed to exactly 0.  Such
// terms are absorbed by the reference's own +1e-100 offsets, so promolecule sums are unchanged;
// the absolute error of a single pro-atom value is < 3.4e-308.
__device__ __forceinline__ bool exp_arg_tiny(double x) {
    return static_cast<unsigned>(__double2hiint(x)) >= 0xC0862000u;
}

// exp(x) for x <= 0, branch-free: Cody-Waite reduction by ln2, degree-11 polynomial (16 FP64 ops).
// GUARD=false drops the underflow guard: only for -700 < x <= 0.
template <bool GUARD = true>
__device__ __forceinline__ double exp_neg_poly(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180559945286e-01, x);
    r = fma(kd, -2.31904681384629956e-17, r);
    double g = fma(HP_EXPG9, r, HP_EXPG8);
    g = fma(g, r, HP_EXPG7);
    g = fma(g, r, HP_EXPG6);
    g = fma(g, r, HP_EXPG5);
    g = fma(g, r, HP_EXPG4);
    g = fma(g, r, HP_EXPG3);
    g = fma(g, r, HP_EXPG2);
    g = fma(g, r, HP_EXPG1);
    g = fma(g, r, HP_EXPG0);
    double p = fma(g, r, 1.0);
    p = fma(p, r, 1.0);
    if (!GUARD) return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    const bool tiny = exp_arg_tiny(x);
    const int hi = tiny ? 0 : __double2hiint(p) + (k << 20);
    const int lo = tiny ? 0 : __double2loint(p);
    return __hiloint2double(hi, lo);
}

}  // namespace hp
