// Per-atom projection and radial pro-atom solves (rows a8, a9, a10 of SURVEY.md section 8a).
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_math.cuh"

namespace hp {

// ---------------------------------------------------------------------------------------------
// Shell tables
// ---------------------------------------------------------------------------------------------
__global__ void table_mbis_kernel(int nshell, const double* __restrict__ propars,
                                  double* __restrict__ A, double* __restrict__ alpha) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nshell) return;
    const double N = propars[2 * k], S = propars[2 * k + 1];
    // mbis.py:286  N * S**3 * exp(-S r) / (8 pi): everything but the exponential
    A[k] = N * (S * S * S) / kEightPi;
    alpha[k] = S;
}

__global__ void table_scaled_kernel(int nshell, const double* __restrict__ c,
                                    const double* __restrict__ norms, double* __restrict__ A) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nshell) A[k] = c[k] * norms[k];
}

__global__ void table_nlis_kernel(int nshell, const double* __restrict__ propars,
                                  const double* __restrict__ inv_gamma, double* __restrict__ A,
                                  double* __restrict__ alpha, double* __restrict__ order) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nshell) return;
    const double N = propars[3 * k], S = propars[3 * k + 1], n = propars[3 * k + 2];
    // nlis.py:325  N * n * S**(3/n) * exp(-S r**n) / (4 pi Gamma(3/n))
    A[k] = N * n * pow(S, 3.0 / n) * inv_gamma[k] / kFourPi;
    alpha[k] = S;
    order[k] = n;
}

// ---------------------------------------------------------------------------------------------
// Spherical average: one warp per radial shell
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shell_project_kernel(int nshell, const int64_t* __restrict__ shell_off, const double* __restrict__ w,
                     const double* __restrict__ rho, const double* __restrict__ atw,
                     const double* __restrict__ shell_r, const double* __restrict__ r2w,
                     double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nshell; s += nwarp) {
        const int64_t lo = shell_off[s], hi = shell_off[s + 1];
        double acc = 0.0;
        for (int64_t p = lo + lane; p < hi; p += 32) acc += (w[p] * rho[p]) * atw[p];
        acc = warp_allsum(acc);
        if (lane == 0) {
            // qc-grid integrate_angular_coordinates + spherical_average: divide the radial factor
            // r^2 w_rad out again, zero at the nucleus, then 1/(4 pi)
            double v = acc / r2w[s];
            if (fabs(shell_r[s]) < 1e-8) v = 0.0;
            out[s] = v / kFourPi;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Real-spherical-harmonic components of w_a*rho on every radial shell (do_density_decomposition,
// core/base.py:637-659 with qc-grid AtomGrid.radial_component_splines): one warp per shell,
//   out[s][lm] = sum_j atw_j f_j Y_lm(Omega_j) / (r_s^2 w_rad_s),   0 at the nucleus,
// Y_lm = sqrt((2l+1)/4pi) R_lm(unit vector), R_lm the Racah-normalised real solid harmonics in
// HORTON-2 order (C_l0, C_l1, S_l1, ...), l <= lmax <= kMaxHarmL.  The harmonics are generated per
// point by the (z, r^2) recursion, m outermost so that only two Legendre-type values are live.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxHarmL = 16;

__global__ void __launch_bounds__(128)
shell_harmonics_kernel(int nshell, int lmax, const int64_t* __restrict__ shell_off,
                       const int* __restrict__ shell_atom, const double* __restrict__ px,
                       const double* __restrict__ py, const double* __restrict__ pz,
                       const double* __restrict__ atom_xyz, const double* __restrict__ w,
                       const double* __restrict__ rho, const double* __restrict__ atw,
                       const double* __restrict__ shell_r, const double* __restrict__ r2w,
                       double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nshell) return;
    const int nlm = (lmax + 1) * (lmax + 1);
    double acc[(kMaxHarmL + 1) * (kMaxHarmL + 1)];
    for (int i = 0; i < nlm; ++i) acc[i] = 0.0;
    const int a = shell_atom[s];
    const double cx = atom_xyz[3 * a], cy = atom_xyz[3 * a + 1], cz = atom_xyz[3 * a + 2];
    const int64_t lo = shell_off[s], hi = shell_off[s + 1];
    for (int64_t p = lo + lane; p < hi; p += 32) {
        const double dx = px[p] - cx, dy = py[p] - cy, dz = pz[p] - cz;
        const double r = sqrt((dx * dx + dy * dy) + dz * dz);
        const double x = r > 0.0 ? dx / r : 0.0, y = r > 0.0 ? dy / r : 0.0, z = r > 0.0 ? dz / r : 0.0;
        const double r2 = x * x + y * y + z * z;
        const double f = (w[p] * rho[p]) * atw[p];
        double A = 1.0, B = 0.0, dfact = 1.0, inv_2m_fact = 1.0;
        for (int m = 0; m <= lmax; ++m) {
            if (m > 0) {
                const double An = x * A - y * B;
                B = x * B + y * A;
                A = An;
                dfact *= double(2 * m - 1);
                inv_2m_fact /= double(2 * m - 1) * double(2 * m);
            }
            double pim2 = 0.0, pim1 = 0.0, ratio = inv_2m_fact;  // (l-m)!/(l+m)! at l = m
            for (int l = m; l <= lmax; ++l) {
                double pi;
                if (l == m) pi = dfact;
                else if (l == m + 1) pi = double(2 * m + 1) * z * pim1;
                else pi = (double(2 * l - 1) * z * pim1 - double(l + m - 1) * r2 * pim2) / double(l - m);
                if (l > m) ratio *= double(l - m) / double(l + m);
                pim2 = pim1;
                pim1 = pi;
                const double ynorm = sqrt(double(2 * l + 1) / kFourPi);
                if (m == 0) {
                    acc[l * l] += f * (ynorm * pi);
                } else {
                    const double nrm = sqrt(2.0 * ratio) * ynorm;
                    acc[l * l + 2 * m - 1] += f * (nrm * pi * A);
                    acc[l * l + 2 * m] += f * (nrm * pi * B);
                }
            }
        }
    }
    const bool nucleus = fabs(shell_r[s]) < 1e-8;
    const double scale = r2w[s];
    for (int i = 0; i < nlm; ++i) {
        const double v = warp_allsum(acc[i]);
        if (lane == 0) out[int64_t(s) * nlm + i] = nucleus ? 0.0 : v / scale;
    }
}

// ---------------------------------------------------------------------------------------------
// MBIS radial fixed point, one warp per atom
// ---------------------------------------------------------------------------------------------
constexpr int kMaxMbisShells = 7;  // periodic table: get_nshell <= 7 (mbis.py:36-46)

__global__ void __launch_bounds__(32)
mbis_radial_kernel(int natom, int atom_base, const int* __restrict__ rad_off, const double* __restrict__ rad_r,
                   const double* __restrict__ rad_w4, const double* __restrict__ sph,
                   const int* __restrict__ par_off, double* __restrict__ propars,
                   const double* __restrict__ pseudo, double threshold, double density_cutoff,
                   int max_inner, double* __restrict__ charges, double* __restrict__ msd,
                   int* __restrict__ niter_out, uint32_t* __restrict__ flags_out) {
    extern __shared__ double smem[];  // [oldpro | pro] each nrad_max
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x;  // global atom index; radial data are rank-local
    const int lane = threadIdx.x;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    const int p0 = par_off[a], K = (par_off[a + 1] - p0) / 2;
    const double* r = rad_r + r0;
    const double* w = rad_w4 + r0;
    const double* rho = sph + r0;
    double* oldpro = smem;

    double N[kMaxMbisShells], S[kMaxMbisShells], N0[kMaxMbisShells], S0[kMaxMbisShells];
#pragma unroll
    for (int k = 0; k < kMaxMbisShells; ++k) {
        N[k] = S[k] = 0.0;
        if (k < K) {
            N[k] = propars[p0 + 2 * k];
            S[k] = propars[p0 + 2 * k + 1];
        }
        N0[k] = N[k];
        S0[k] = S[k];
    }

    // pop = sum weights * rho  (mbis.py:122); also the pseudo-population of mbis.py:201
    double pop = 0.0;
    for (int i = lane; i < nrad; i += 32) pop += w[i] * rho[i];
    pop = warp_allsum(pop);

    uint32_t flags = HP_SOLVE_NOT_CONVERGED;
    int it = 0;
    for (; it < max_inner; ++it) {
        double m0[kMaxMbisShells], m1[kMaxMbisShells];
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) m0[k] = m1[k] = 0.0;
        double chg = 0.0;
        for (int i = lane; i < nrad; i += 32) {
            const double ri = r[i];
            double term[kMaxMbisShells];
            double pro = 0.0;
#pragma unroll
            for (int k = 0; k < kMaxMbisShells; ++k) {
                term[k] = 0.0;
                if (k < K) {
                    // mbis.py:128
                    term[k] = N[k] * (S[k] * S[k] * S[k]) * exp(-S[k] * ri) / kEightPi;
                    pro += term[k];
                }
            }
            const double rh = rho[i];
            const bool sick = (rh < density_cutoff) || (pro < density_cutoff);
            const double ratio = sick ? 0.0 : rh / pro;
#pragma unroll
            for (int k = 0; k < kMaxMbisShells; ++k) {
                if (k < K) {
                    const double tr = term[k] * ratio;
                    m0[k] += w[i] * tr;          // mbis.py:143
                    m1[k] += w[i] * tr * ri;     // mbis.py:144
                }
            }
            if (it > 0) {
                const double e = oldpro[i] - pro;
                chg += w[i] * e * e;             // mbis.py:151-152
            }
            oldpro[i] = pro;
        }
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) {
            if (k < K) {
                const double a0 = warp_allsum(m0[k]);
                const double a1 = warp_allsum(m1[k]);
                N[k] = a0;                       // mbis.py:145
                S[k] = 3.0 * a0 / a1;            // mbis.py:146
            }
        }
        const double change = (it == 0) ? 1e100 : sqrt(warp_allsum(chg));
        if (change < threshold) {
            flags &= ~HP_SOLVE_NOT_CONVERGED;
            ++it;
            break;
        }
    }

    double nsum = 0.0;
    bool finite = true;
#pragma unroll
    for (int k = 0; k < kMaxMbisShells; ++k) {
        if (k < K) {
            nsum += N[k];
            finite = finite && isfinite(N[k]) && isfinite(S[k]);
        }
    }
    // mbis.py:157 np.isclose(pop, sum N, atol=1e-4) -> |a-b| <= atol + rtol*|b|, rtol = 1e-5
    if (!(fabs(pop - nsum) <= 1e-4 + 1e-5 * fabs(nsum))) flags |= HP_SOLVE_POP_MISMATCH;
    if (!finite) flags |= HP_SOLVE_NONFINITE;

    // this atom's contribution to compute_change (core/iterstock.py:36-44, mbis.py:256-260)
    double dev = 0.0;
    for (int i = lane; i < nrad; i += 32) {
        const double ri = r[i];
        double ynew = 0.0, yold = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) {
            if (k < K) {
                ynew += N[k] * (S[k] * S[k] * S[k]) * exp(-S[k] * ri) / kEightPi;
                yold += N0[k] * (S0[k] * S0[k] * S0[k]) * exp(-S0[k] * ri) / kEightPi;
            }
        }
        const double d = ynew - yold;
        dev += w[i] * d * d;
    }
    dev = warp_allsum(dev);

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) {
            if (k < K) {
                propars[p0 + 2 * k] = N[k];
                propars[p0 + 2 * k + 1] = S[k];
            }
        }
        charges[a] = pseudo[a] - pop;  // mbis.py:203
        msd[a] = dev;
        niter_out[a] = it;
        flags_out[a] = flags;
    }
}

// ---------------------------------------------------------------------------------------------
// MBIS / NLIS radial fixed point, one BLOCK of 128 threads per atom (<= 256 radial points): the same
// arithmetic per point as the one-warp kernels above/below, but a thread owns at most two radial points
// instead of five to eight, and the 2K+1 sums of an inner iteration are reduced by warp butterflies +
// one barrier.  The inner loop is a chain of ~30-100 strictly sequential steps per outer iteration, each
// bounded by the latency of exp + two divisions; with one warp per atom that chain was 65 % of the GPU
// time of a small system (round-1 smoke profile).  NLIS = true: shells (N, S, n), nlis.py:99-194.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pow_order(double r, double n);
__device__ __forceinline__ double rcp_newton(double x);

constexpr int kRadThreads = 128;
constexpr int kRadWarps = kRadThreads / 32;
constexpr int kRadOwn = 2;  // radial points per thread

template <bool NLIS, int KMAX>
__global__ void __launch_bounds__(kRadThreads)
shell_fixed_point_block_kernel(int natom, int atom_base, const int* __restrict__ rad_off,
                               const double* __restrict__ rad_r, const double* __restrict__ rad_w4,
                               const double* __restrict__ sph, const int* __restrict__ par_off,
                               double* __restrict__ propars, const int* __restrict__ shell_off,
                               const double* __restrict__ inv_gamma, const double* __restrict__ pseudo,
                               double threshold, double density_cutoff, int max_inner,
                               double* __restrict__ charges, double* __restrict__ msd,
                               int* __restrict__ niter_out, uint32_t* __restrict__ flags_out) {
    constexpr int NV = 2 * KMAX + 1;  // (m0, m1) per shell | change term (also used for pop / dev)
    __shared__ double s_part[2][kRadWarps][NV];
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    constexpr int PW = NLIS ? 3 : 2;
    const int p0 = par_off[a], K = (par_off[a + 1] - p0) / PW;
    const double* ig = NLIS ? inv_gamma + shell_off[a] : nullptr;

    double N[KMAX], S[KMAX], n[KMAX], G[KMAX], N0[KMAX],
        S0[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        N[k] = S[k] = 0.0;
        n[k] = G[k] = 1.0;
        if (k < K) {
            N[k] = propars[p0 + PW * k];
            S[k] = propars[p0 + PW * k + 1];
            if (NLIS) {
                n[k] = propars[p0 + 3 * k + 2];
                G[k] = ig[k];
            }
        }
        N0[k] = N[k];
        S0[k] = S[k];
    }
    double ri[kRadOwn], wi[kRadOwn], rhoi[kRadOwn], oldpro[kRadOwn];
    bool live[kRadOwn];
#pragma unroll
    for (int q = 0; q < kRadOwn; ++q) {
        const int i = tid + q * kRadThreads;
        live[q] = i < nrad;
        ri[q] = live[q] ? rad_r[r0 + i] : 1.0;
        wi[q] = live[q] ? rad_w4[r0 + i] : 0.0;
        rhoi[q] = live[q] ? sph[r0 + i] : 0.0;
        oldpro[q] = 0.0;
    }
    // all-reduce of per-thread values: butterflies, one barrier, fixed summation order.  Slots 0..2K-1 hold
    // (m0, m1) per shell, slot CHG the change term; every loop is unrolled so that `v` stays in registers.
    constexpr int CHG = 2 * KMAX;
    int parity = 0;
    auto block_allsum = [&](double (&v)[NV], int nshell2) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if (j < nshell2 || j == CHG) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
                if (lane == 0) s_part[parity][warp][j] = v[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if (j < nshell2 || j == CHG) {
                double t = s_part[parity][0][j];
#pragma unroll
                for (int wq = 1; wq < kRadWarps; ++wq) t += s_part[parity][wq][j];
                v[j] = t;
            }
        }
        parity ^= 1;  // the next reduction writes the other buffer: no second barrier needed
    };

    // MBIS: one shell's pro-atom term N S^3 exp(-S r) / (8 pi) (mbis.py:128).  NLIS: the normalised shell
    // function WITHOUT its population, n S^(3/n) exp(-S r^n) / (4 pi Gamma(3/n)) (nlis.py:147).
    auto shell_term = [&](int k, double Nk, double Sk, double r, double& rn) -> double {
        if (!NLIS) {
            rn = r;
            // exp_neg_poly: the 16-operation inline exponential of the grid kernels (<= 0.67 ulp) instead of
            // the library call -- this term sits on the sequential chain of the fixed point
            return Nk * (Sk * Sk * Sk) * exp_neg_poly(-Sk * r) / kEightPi;
        }
        rn = pow_order(r, n[k]);
        return n[k] * pow(Sk, 3.0 / n[k]) * exp(-Sk * rn) * G[k] / kFourPi;
    };

    double red[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) red[j] = 0.0;
#pragma unroll
    for (int q = 0; q < kRadOwn; ++q) red[CHG] += wi[q] * rhoi[q];  // mbis.py:122, :201
    block_allsum(red, 0);
    const double pop = red[CHG];

    uint32_t flags = HP_SOLVE_NOT_CONVERGED;
    int it = 0;
    for (; it < max_inner; ++it) {
#pragma unroll
        for (int j = 0; j < NV; ++j) red[j] = 0.0;
#pragma unroll
        for (int q = 0; q < kRadOwn; ++q) {
            double term[KMAX], rn[KMAX];
            double pro = 0.0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                term[k] = 0.0;
                rn[k] = 0.0;
                if (k < K) {
                    term[k] = shell_term(k, N[k], S[k], ri[q], rn[k]);
                    pro += NLIS ? term[k] * N[k] : term[k];  // nlis.py:150
                }
            }
            const bool sick = (rhoi[q] < density_cutoff) || (pro < density_cutoff);
            const double ratio = sick ? 0.0 : rhoi[q] * rcp_newton(pro);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    const double tr = term[k] * ratio;
                    red[2 * k] += NLIS ? wi[q] * (tr * N[k]) : wi[q] * tr;  // nlis.py:166 / mbis.py:143
                    red[2 * k + 1] += wi[q] * tr * rn[k];                   // nlis.py:167 / mbis.py:144
                }
            }
            if (it > 0) {
                const double e = oldpro[q] - pro;
                red[CHG] += wi[q] * e * e;  // mbis.py:151-152
            }
            oldpro[q] = pro;
        }
        block_allsum(red, 2 * K);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                const double a0 = red[2 * k], a1 = red[2 * k + 1];
                if (!NLIS) {
                    N[k] = a0;                 // mbis.py:145
                    S[k] = 3.0 * a0 / a1;      // mbis.py:146
                } else {
                    N[k] = a0;
                    // nlis.py:170-176: np.isclose(m1, 0) -> |m1| <= 1e-8
                    S[k] = (fabs(a1) <= 1e-8) ? 1e-5 : 3.0 / (a1 * n[k]);
                }
            }
        }
        const double change = (it == 0) ? 1e100 : sqrt(red[CHG]);
        if (change < threshold) {
            flags &= ~HP_SOLVE_NOT_CONVERGED;
            ++it;
            break;
        }
    }

    double nsum = 0.0;
    bool finite = true;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            nsum += N[k];
            finite = finite && isfinite(N[k]) && isfinite(S[k]);
        }
    }
    if (!(fabs(pop - nsum) <= 1e-4 + 1e-5 * fabs(nsum))) flags |= HP_SOLVE_POP_MISMATCH;  // mbis.py:157
    if (!finite) flags |= HP_SOLVE_NONFINITE;

    // this atom's contribution to compute_change (core/iterstock.py:36-44)
    red[CHG] = 0.0;
#pragma unroll
    for (int q = 0; q < kRadOwn; ++q) {
        double ynew = 0.0, yold = 0.0, rn;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                if (!NLIS) {
                    ynew += shell_term(k, N[k], S[k], ri[q], rn);
                    yold += shell_term(k, N0[k], S0[k], ri[q], rn);
                } else {  // nlis.py:296-299
                    rn = pow_order(ri[q], n[k]);
                    ynew += N[k] * n[k] * pow(S[k], 3.0 / n[k]) * exp(-S[k] * rn) * G[k] / kFourPi;
                    yold += N0[k] * n[k] * pow(S0[k], 3.0 / n[k]) * exp(-S0[k] * rn) * G[k] / kFourPi;
                }
            }
        }
        const double d = ynew - yold;
        red[CHG] += wi[q] * d * d;
    }
    block_allsum(red, 0);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                propars[p0 + PW * k] = N[k];
                propars[p0 + PW * k + 1] = S[k];
            }
        }
        charges[a] = pseudo[a] - pop;  // mbis.py:203
        msd[a] = red[CHG];
        niter_out[a] = it;
        flags_out[a] = flags;
    }
}

// ---------------------------------------------------------------------------------------------
// NLIS / GMBIS radial fixed point (nlis.py:99-194), one warp per atom.  Shells are (N, S, n) with
// n fixed; inv_gamma[k] = 1/Gamma(3/n_k) comes from the host.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pow_order(double r, double n) {
    return (n == 1.0) ? r : ((n == 2.0) ? r * r : pow(r, n));
}

__global__ void __launch_bounds__(32)
nlis_radial_kernel(int natom, int atom_base, const int* __restrict__ rad_off,
                   const double* __restrict__ rad_r, const double* __restrict__ rad_w4,
                   const double* __restrict__ sph, const int* __restrict__ par_off,
                   double* __restrict__ propars, const int* __restrict__ shell_off,
                   const double* __restrict__ inv_gamma, const double* __restrict__ pseudo,
                   double threshold, double density_cutoff, int max_inner,
                   double* __restrict__ charges, double* __restrict__ msd,
                   int* __restrict__ niter_out, uint32_t* __restrict__ flags_out) {
    extern __shared__ double smem[];
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x;
    const int lane = threadIdx.x;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    const int p0 = par_off[a], K = (par_off[a + 1] - p0) / 3;
    const double* r = rad_r + r0;
    const double* w = rad_w4 + r0;
    const double* rho = sph + r0;
    const double* ig = inv_gamma + shell_off[a];
    double* oldpro = smem;

    double N[kMaxMbisShells], S[kMaxMbisShells], n[kMaxMbisShells], N0[kMaxMbisShells],
        S0[kMaxMbisShells], G[kMaxMbisShells];
#pragma unroll
    for (int k = 0; k < kMaxMbisShells; ++k) {
        N[k] = S[k] = 0.0;
        n[k] = G[k] = 1.0;
        if (k < K) {
            N[k] = propars[p0 + 3 * k];
            S[k] = propars[p0 + 3 * k + 1];
            n[k] = propars[p0 + 3 * k + 2];
            G[k] = ig[k];
        }
        N0[k] = N[k];
        S0[k] = S[k];
    }
    double pop = 0.0;
    for (int i = lane; i < nrad; i += 32) pop += w[i] * rho[i];
    pop = warp_allsum(pop);

    uint32_t flags = HP_SOLVE_NOT_CONVERGED;
    int it = 0;
    for (; it < max_inner; ++it) {
        double m0[kMaxMbisShells], m1[kMaxMbisShells];
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) m0[k] = m1[k] = 0.0;
        double chg = 0.0;
        for (int i = lane; i < nrad; i += 32) {
            const double ri = r[i];
            double g[kMaxMbisShells], rn[kMaxMbisShells];
            double pro = 0.0;
#pragma unroll
            for (int k = 0; k < kMaxMbisShells; ++k) {
                g[k] = rn[k] = 0.0;
                if (k < K) {
                    rn[k] = pow_order(ri, n[k]);
                    // nlis.py:147  n S^(3/n) exp(-S r^n) / (4 pi Gamma(3/n))
                    g[k] = n[k] * pow(S[k], 3.0 / n[k]) * exp(-S[k] * rn[k]) * G[k] / kFourPi;
                    pro += g[k] * N[k];  // nlis.py:150
                }
            }
            const double rh = rho[i];
            const bool sick = (rh < density_cutoff) || (pro < density_cutoff);
            const double ratio = sick ? 0.0 : rh / pro;
#pragma unroll
            for (int k = 0; k < kMaxMbisShells; ++k) {
                if (k < K) {
                    const double tr = g[k] * ratio;
                    m0[k] += w[i] * (tr * N[k]);   // nlis.py:166
                    m1[k] += w[i] * tr * rn[k];    // nlis.py:167
                }
            }
            if (it > 0) {
                const double e = oldpro[i] - pro;
                chg += w[i] * e * e;
            }
            oldpro[i] = pro;
        }
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) {
            if (k < K) {
                const double a0 = warp_allsum(m0[k]);
                const double a1 = warp_allsum(m1[k]);
                N[k] = a0;
                // nlis.py:170-176: np.isclose(m1, 0) -> |m1| <= 1e-8
                S[k] = (fabs(a1) <= 1e-8) ? 1e-5 : 3.0 / (a1 * n[k]);
            }
        }
        const double change = (it == 0) ? 1e100 : sqrt(warp_allsum(chg));
        if (change < threshold) {
            flags &= ~HP_SOLVE_NOT_CONVERGED;
            ++it;
            break;
        }
    }
    double nsum = 0.0;
    bool finite = true;
#pragma unroll
    for (int k = 0; k < kMaxMbisShells; ++k) {
        if (k < K) {
            nsum += N[k];
            finite = finite && isfinite(N[k]) && isfinite(S[k]);
        }
    }
    if (!(fabs(pop - nsum) <= 1e-4 + 1e-5 * fabs(nsum))) flags |= HP_SOLVE_POP_MISMATCH;
    if (!finite) flags |= HP_SOLVE_NONFINITE;

    // compute_change term (core/iterstock.py:36-44 with nlis.py:296-299)
    double dev = 0.0;
    for (int i = lane; i < nrad; i += 32) {
        const double ri = r[i];
        double ynew = 0.0, yold = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) {
            if (k < K) {
                const double rn = pow_order(ri, n[k]);
                ynew += N[k] * n[k] * pow(S[k], 3.0 / n[k]) * exp(-S[k] * rn) * G[k] / kFourPi;
                yold += N0[k] * n[k] * pow(S0[k], 3.0 / n[k]) * exp(-S0[k] * rn) * G[k] / kFourPi;
            }
        }
        const double d = ynew - yold;
        dev += w[i] * d * d;
    }
    dev = warp_allsum(dev);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kMaxMbisShells; ++k) {
            if (k < K) {
                propars[p0 + 3 * k] = N[k];
                propars[p0 + 3 * k + 1] = S[k];
            }
        }
        charges[a] = pseudo[a] - pop;
        msd[a] = dev;
        niter_out[a] = it;
        flags_out[a] = flags;
    }
}

// ---------------------------------------------------------------------------------------------
// aLISA self-consistent update (alisa.py:193-291 `solver_sc`, :294-353 `solver_sc_1_iter`,
// utils.py:198-252 `compute_quantities`), one warp per atom.  Basis functions on the radial grid
// (K x nrad per atom, evaluated once on the host exactly as gisa.py:91-106 does) are read through
// L1; the coefficient vector lives in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
lisa_sc_radial_kernel(int natom, int atom_base, const int* __restrict__ rad_off,
                      const double* __restrict__ rad_w4, const double* __restrict__ sph,
                      const int* __restrict__ par_off, double* __restrict__ propars,
                      const int64_t* __restrict__ bs_off, const double* __restrict__ bs,
                      const double* __restrict__ pseudo, double threshold, double density_cutoff,
                      double population_cutoff, int max_inner, int single_update, int nrad_max,
                      double* __restrict__ charges, double* __restrict__ msd,
                      int* __restrict__ niter_out, uint32_t* __restrict__ flags_out) {
    extern __shared__ double smem[];  // oldpro[nrad_max] | ratio[nrad_max] | c[K] | c0[K]
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x;
    const int lane = threadIdx.x;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    const int p0 = par_off[a], K = par_off[a + 1] - p0;
    const double* w = rad_w4 + r0;
    const double* rho = sph + r0;
    const double* g = bs + bs_off[blockIdx.x];  // g[k * nrad + i]
    double* oldpro = smem;
    double* ratio = smem + nrad_max;
    double* c = smem + 2 * nrad_max;
    double* c0 = c + K;
    for (int k = lane; k < K; k += 32) c[k] = c0[k] = propars[p0 + k];
    __syncwarp();

    double pop = 0.0;
    for (int i = lane; i < nrad; i += 32) pop += w[i] * rho[i];
    pop = warp_allsum(pop);

    uint32_t flags = single_update ? 0u : HP_SOLVE_NOT_CONVERGED;
    int it = 0;
    for (; it < max_inner; ++it) {
        double chg = 0.0;
        for (int i = lane; i < nrad; i += 32) {
            double pro = 0.0;
            for (int k = 0; k < K; ++k) pro += g[k * nrad + i] * c[k];  // utils.py:236-237
            const double rh = rho[i];
            const bool sick = (rh < density_cutoff) || (pro < density_cutoff);
            ratio[i] = sick ? 0.0 : rh / pro;
            if (it > 0) {
                const double e = oldpro[i] - pro;
                chg += w[i] * e * e;  // alisa.py:273-274
            }
            oldpro[i] = pro;
        }
        __syncwarp();
        for (int k = 0; k < K; ++k) {
            const double ck = c[k];
            double s = 0.0;
            for (int i = lane; i < nrad; i += 32) s += w[i] * ((g[k * nrad + i] * ck) * ratio[i]);
            s = warp_allsum(s);  // alisa.py:268
            __syncwarp();
            if (lane == 0) c[k] = s;
        }
        __syncwarp();
        if (single_update) {
            ++it;
            break;
        }
        const double change = (it == 0) ? 1e100 : sqrt(warp_allsum(chg));
        if (change < threshold) {
            flags &= ~HP_SOLVE_NOT_CONVERGED;
            ++it;
            break;
        }
    }
    double csum = 0.0;
    bool finite = true;
    for (int k = 0; k < K; ++k) {
        csum += c[k];
        finite = finite && isfinite(c[k]);
    }
    // check_pars_population, utils.py:434: only reached on convergence in the reference
    if (!single_update && !(flags & HP_SOLVE_NOT_CONVERGED) && fabs(csum - pop) > population_cutoff)
        flags |= HP_SOLVE_POP_MISMATCH;
    if (!finite) flags |= HP_SOLVE_NONFINITE;

    double dev = 0.0;
    for (int i = lane; i < nrad; i += 32) {
        double ynew = 0.0, yold = 0.0;
        for (int k = 0; k < K; ++k) {
            ynew += c[k] * g[k * nrad + i];
            yold += c0[k] * g[k * nrad + i];
        }
        const double d = ynew - yold;
        dev += w[i] * d * d;
    }
    dev = warp_allsum(dev);
    for (int k = lane; k < K; k += 32) propars[p0 + k] = c[k];
    if (lane == 0) {
        charges[a] = pseudo[a] - pop;  // gisa.py:315-318
        msd[a] = dev;
        niter_out[a] = it;
        flags_out[a] = flags;
    }
}

// ---------------------------------------------------------------------------------------------
// aLISA self-consistent update, one BLOCK of kScWarps warps per atom (the inner fixed point runs
// ~10^4 strictly sequential steps for the Slater basis, so the step LATENCY is what counts).
//
// Registers: warp v owns the shells k = v, v + kScWarps, ... (at most KPW of them) for ALL radial
// points; lane l holds the points i = l + 32 j (j < NPL).  g[k][i] of the owned shells, the
// coefficients c[k] and nothing else live in registers for the whole solve -- no global or
// shared-memory traffic for the K x nrad table inside the loop.
// One step = two barriers:
//   1. every warp: partial pro-atom sum over its shells for all points        -> s_part[v][i]
//   2. the thread that owns point i (i = tid, tid + 128): pro = sum_v s_part[v][i] (fixed order),
//      ratio = rho / pro (masked, utils.py:238-245), w * ratio and the change term w (oldpro - pro)^2
//                                                                            -> s_rw[i], s_chg[i]
//   3. every warp: c_k <- c_k * sum_i g_k(i) (w ratio)(i) for its shells (alisa.py:268) and,
//      redundantly, the change (alisa.py:273-274): all warps read the same numbers in the same order,
//      so the stopping decision is uniform without another barrier.
// Arithmetic differs from the one-warp kernel only in the association of the products
// (c_k * sum g (w ratio) instead of sum w ((g c_k) ratio)) and in the order of the K-term sum.
// ---------------------------------------------------------------------------------------------
constexpr int kScWarps = 4;
constexpr int kScThreads = kScWarps * 32;

// 1 / x for a normal, positive x: MUFU.RCP64H seed and two Newton steps (relative error of the result below
// 2^-52 before the caller's multiplication: rho * (1 / pro) differs from the correctly rounded quotient by at
// most one unit in the last place).  Six dependent operations instead of the ~ten of a full division: the
// inner fixed point is a chain of ~1e4 sequential steps and this quotient sits on it.  x below 1e-15 never
// gets here (masked by density_cutoff).
__device__ __forceinline__ double rcp_newton(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

template <int KPW, int NPL>
__global__ void __launch_bounds__(kScThreads)
lisa_sc_block_kernel(int natom, int atom_base, const int* __restrict__ rad_off,
                     const double* __restrict__ rad_w4, const double* __restrict__ sph,
                     const int* __restrict__ par_off, double* __restrict__ propars,
                     const int64_t* __restrict__ bs_off, const double* __restrict__ bs,
                     const double* __restrict__ pseudo, double threshold, double density_cutoff,
                     double population_cutoff, int max_inner, int single_update,
                     double* __restrict__ charges, double* __restrict__ msd,
                     int* __restrict__ niter_out, uint32_t* __restrict__ flags_out) {
    constexpr int NRP = NPL * 32;            // padded radial size
    constexpr int NOWN = (NRP + kScThreads - 1) / kScThreads;  // points owned per thread in step 2
    __shared__ double s_part[kScWarps][NRP];
    __shared__ double s_rw[NRP];
    __shared__ double s_chg[NRP];
    __shared__ double s_red[32];
    __shared__ double s_c[KPW * kScWarps];
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    const int p0 = par_off[a], K = par_off[a + 1] - p0;
    const double* w = rad_w4 + r0;
    const double* rho = sph + r0;
    const double* gsrc = bs + bs_off[blockIdx.x];  // g[k * nrad + i]

    double g[KPW][NPL], c[KPW], c0[KPW];
#pragma unroll
    for (int kk = 0; kk < KPW; ++kk) {
        const int k = warp + kk * kScWarps;
        c[kk] = c0[kk] = (k < K) ? propars[p0 + k] : 0.0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int i = lane + 32 * j;
            g[kk][j] = (k < K && i < nrad) ? gsrc[int64_t(k) * nrad + i] : 0.0;
        }
    }
    double wo[NOWN], ro[NOWN], oldpro[NOWN];
    double pop = 0.0;
#pragma unroll
    for (int q = 0; q < NOWN; ++q) {
        const int i = tid + q * kScThreads;
        wo[q] = (i < nrad) ? w[i] : 0.0;
        ro[q] = (i < nrad) ? rho[i] : 0.0;
        oldpro[q] = 0.0;
        pop += wo[q] * ro[q];
    }
    pop = block_sum(pop, s_red);  // thread 0 only
    if (tid == 0) s_red[0] = pop;
    __syncthreads();
    pop = s_red[0];
    __syncthreads();

    uint32_t flags = single_update ? 0u : HP_SOLVE_NOT_CONVERGED;
    int it = 0;
    for (; it < max_inner; ++it) {
        // 1. partial pro-atom sums of this warp's shells
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            double pa = 0.0, pb = 0.0;  // two chains: the step latency is what counts here
#pragma unroll
            for (int kk = 0; kk < KPW; kk += 2) {
                pa = fma(g[kk][j], c[kk], pa);
                if (kk + 1 < KPW) pb = fma(g[kk + 1][j], c[kk + 1], pb);
            }
            s_part[warp][lane + 32 * j] = pa + pb;
        }
        __syncthreads();
        // 2. point owners: pro-atom, masked ratio, change term
#pragma unroll
        for (int q = 0; q < NOWN; ++q) {
            const int i = tid + q * kScThreads;
            if (i < NRP) {
                static_assert(kScWarps == 4, "pairwise sum below is written for four warps");
                const double pro = (s_part[0][i] + s_part[1][i]) + (s_part[2][i] + s_part[3][i]);
                const bool sick = (ro[q] < density_cutoff) || (pro < density_cutoff);
                const double ratio = sick ? 0.0 : ro[q] * rcp_newton(pro);
                const double e = oldpro[q] - pro;
                s_rw[i] = wo[q] * ratio;
                s_chg[i] = (it > 0) ? wo[q] * e * e : 0.0;
                oldpro[q] = pro;
            }
        }
        __syncthreads();
        // 3. coefficient update of this warp's shells + the (redundant) change
        double rw[NPL], chg = 0.0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            rw[j] = s_rw[lane + 32 * j];
            chg += s_chg[lane + 32 * j];
        }
        double sums[KPW];
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) {
            double sa = 0.0, sb = 0.0;
#pragma unroll
            for (int j = 0; j < NPL; j += 2) {
                sa = fma(g[kk][j], rw[j], sa);
                if (j + 1 < NPL) sb = fma(g[kk][j + 1], rw[j + 1], sb);
            }
            sums[kk] = sa + sb;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int kk = 0; kk < KPW; ++kk) sums[kk] += __shfl_xor_sync(0xffffffffu, sums[kk], off);
            chg += __shfl_xor_sync(0xffffffffu, chg, off);
        }
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) c[kk] *= sums[kk];
        if (single_update) {
            ++it;
            break;
        }
        const double change = (it == 0) ? 1e100 : sqrt(chg);
        if (change < threshold) {
            flags &= ~HP_SOLVE_NOT_CONVERGED;
            ++it;
            break;
        }
    }

    // this atom's term of compute_change (core/iterstock.py:36-44) and the population check
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        double pn = 0.0, po = 0.0;
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) {
            pn = fma(g[kk][j], c[kk], pn);
            po = fma(g[kk][j], c0[kk], po);
        }
        s_part[warp][lane + 32 * j] = pn - po;
    }
    if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) {
            const int k = warp + kk * kScWarps;
            s_c[k] = c[kk];
            if (k < K) propars[p0 + k] = c[kk];
        }
    }
    __syncthreads();
    double dev = 0.0;
#pragma unroll
    for (int q = 0; q < NOWN; ++q) {
        const int i = tid + q * kScThreads;
        if (i < NRP) {
            const double d = (s_part[0][i] + s_part[1][i]) + (s_part[2][i] + s_part[3][i]);
            dev += wo[q] * d * d;
        }
    }
    dev = block_sum(dev, s_red);
    if (tid == 0) {
        double csum = 0.0;
        bool finite = true;
        for (int k = 0; k < K; ++k) {
            csum += s_c[k];
            finite = finite && isfinite(s_c[k]);
        }
        // check_pars_population, utils.py:434: only reached on convergence in the reference
        if (!single_update && !(flags & HP_SOLVE_NOT_CONVERGED) && fabs(csum - pop) > population_cutoff)
            flags |= HP_SOLVE_POP_MISMATCH;
        if (!finite) flags |= HP_SOLVE_NONFINITE;
        charges[a] = pseudo[a] - pop;  // gisa.py:315-318
        msd[a] = dev;
        niter_out[a] = it;
        flags_out[a] = flags;
    }
}

// ---------------------------------------------------------------------------------------------
// Row a13: multipole moments of w_a*rho about R_a on atom a's own atomic grid (core/base.py:329-402
// with qc-grid Grid.moments): Cartesian monomials in HORTON order, real regular solid harmonics
// (Racah normalisation, order C_l0 C_l1 S_l1 C_l2 S_l2 ...), radial moments r^n.  One block per
// atom; out row = [ncart | npure | nrad] raw integrals (signs/offsets are applied by the caller).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxL = 4;
constexpr int kMaxMom = 35 + 25 + 5;

__device__ __forceinline__ int fill_moment_basis(int lmax, double x, double y, double z, double* b) {
    double xp[kMaxL + 1], yp[kMaxL + 1], zp[kMaxL + 1];
    xp[0] = yp[0] = zp[0] = 1.0;
    for (int i = 1; i <= lmax; ++i) {
        xp[i] = xp[i - 1] * x;
        yp[i] = yp[i - 1] * y;
        zp[i] = zp[i - 1] * z;
    }
    int n = 0;
    for (int l = 0; l <= lmax; ++l)
        for (int nx = l; nx >= 0; --nx)
            for (int ny = l - nx; ny >= 0; --ny) b[n++] = xp[nx] * yp[ny] * zp[l - nx - ny];
    // solid harmonics: A_m + i B_m = (x + i y)^m, Pi_l^m(z, r^2) by the Legendre-type recursion
    const double r2 = x * x + y * y + z * z;
    double A[kMaxL + 1], B[kMaxL + 1];
    A[0] = 1.0;
    B[0] = 0.0;
    for (int m = 1; m <= lmax; ++m) {
        A[m] = x * A[m - 1] - y * B[m - 1];
        B[m] = x * B[m - 1] + y * A[m - 1];
    }
    double Pi[kMaxL + 1][kMaxL + 1];
    for (int m = 0; m <= lmax; ++m) {
        double df = 1.0;
        for (int k = 1; k < 2 * m; k += 2) df *= k;
        Pi[m][m] = df;
        if (m + 1 <= lmax) Pi[m + 1][m] = (2 * m + 1) * z * Pi[m][m];
        for (int l = m + 2; l <= lmax; ++l)
            Pi[l][m] = ((2 * l - 1) * z * Pi[l - 1][m] - (l + m - 1) * r2 * Pi[l - 2][m]) / (l - m);
    }
    for (int l = 0; l <= lmax; ++l) {
        b[n++] = Pi[l][0];
        double ratio = 1.0;  // (l-m)!/(l+m)!
        for (int m = 1; m <= l; ++m) {
            ratio /= double(l + m) * double(l - m + 1);
            const double norm = sqrt(2.0 * ratio);
            b[n++] = norm * Pi[l][m] * A[m];
            b[n++] = norm * Pi[l][m] * B[m];
        }
    }
    const double r = sqrt(r2);
    double rp = 1.0;
    for (int k = 0; k <= lmax; ++k) {
        b[n++] = rp;
        rp *= r;
    }
    return n;
}

__global__ void __launch_bounds__(128)
atom_moments_kernel(int natom, int atom_base, int lmax, const int64_t* __restrict__ seg_off,
                    const double* __restrict__ px, const double* __restrict__ py,
                    const double* __restrict__ pz, const double* __restrict__ atw,
                    const double* __restrict__ at_w, const double* __restrict__ dens,
                    const double* __restrict__ atom_xyz, int nmom, double* __restrict__ out) {
    __shared__ double red[32];
    const int la = blockIdx.x;
    if (la >= natom) return;
    const int a = atom_base + la;
    const double cx = atom_xyz[3 * a], cy = atom_xyz[3 * a + 1], cz = atom_xyz[3 * a + 2];
    double acc[kMaxMom];
    for (int i = 0; i < kMaxMom; ++i) acc[i] = 0.0;
    for (int64_t p = seg_off[la] + threadIdx.x; p < seg_off[la + 1]; p += blockDim.x) {
        double b[kMaxMom];
        fill_moment_basis(lmax, px[p] - cx, py[p] - cy, pz[p] - cz, b);
        const double f = atw[p] * (dens[p] * at_w[p]);
        for (int i = 0; i < nmom; ++i) acc[i] = fma(b[i], f, acc[i]);
    }
    for (int i = 0; i < nmom; ++i) {
        const double t = block_sum(acc[i], red);
        if (threadIdx.x == 0) out[int64_t(a) * nmom + i] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// End-of-iteration scalars in a fixed summation order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
finish_iteration_kernel(int npartial, const double* __restrict__ partials, int natom,
                        const double* __restrict__ msd, double* __restrict__ out2) {
    __shared__ double red[32];
    double e = 0.0;
    if (partials)
        for (int i = threadIdx.x; i < npartial; i += blockDim.x) e += partials[i];
    e = block_sum(e, red);
    double m = 0.0;
    for (int i = threadIdx.x; i < natom; i += blockDim.x) m += msd[i];
    m = block_sum(m, red);
    if (threadIdx.x == 0) {
        out2[0] = sqrt(m);  // core/iterstock.py:45
        out2[1] = e;
    }
}

__global__ void __launch_bounds__(256)
sum_partials_kernel(int n, const double* __restrict__ partials, double* __restrict__ out) {
    __shared__ double red[32];
    double e = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) e += partials[i];
    e = block_sum(e, red);
    if (threadIdx.x == 0) out[0] = e;
}

// out[s] = sum_{p in segment s} w[p] * f[p] * (g ? g[p] : 1): one block per segment, fixed order
__global__ void __launch_bounds__(256)
segment_integrate_kernel(int nseg, const int64_t* __restrict__ seg_off, const double* __restrict__ w,
                         const double* __restrict__ f, const double* __restrict__ g,
                         double* __restrict__ out) {
    __shared__ double red[32];
    for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
        const int64_t lo = seg_off[s], hi = seg_off[s + 1];
        double acc = 0.0;
        for (int64_t p = lo + threadIdx.x; p < hi; p += blockDim.x)
            acc += g ? (w[p] * f[p]) * g[p] : w[p] * f[p];
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) out[s] = acc;
    }
}

}  // namespace hp

using namespace hp;

extern "C" int hp_sum_partials(int32_t n, const double* partials, double* out1, void* stream) {
    HP_REQUIRE(n > 0 && partials && out1, "bad arguments");
    sum_partials_kernel<<<1, 256, 0, as_stream(stream)>>>(n, partials, out1);
    HP_LAUNCH_CHECK("sum_partials_kernel");
    return HP_OK;
}

extern "C" int hp_segment_integrate(int32_t nseg, const int64_t* seg_offsets, const double* w,
                                    const double* f, const double* g, double* out, void* stream) {
    HP_REQUIRE(nseg >= 0, "bad sizes");
    if (nseg == 0) return HP_OK;
    HP_REQUIRE(seg_offsets && w && f && out, "null input");
    int blocks = nseg < sm_count() * 8 ? nseg : sm_count() * 8;
    segment_integrate_kernel<<<blocks, 256, 0, as_stream(stream)>>>(nseg, seg_offsets, w, f, g, out);
    HP_LAUNCH_CHECK("segment_integrate_kernel");
    return HP_OK;
}

extern "C" int hp_table_mbis(int32_t nshell, const double* propars, double* shell_A,
                             double* shell_alpha, void* stream) {
    HP_REQUIRE(nshell > 0 && propars && shell_A && shell_alpha, "bad arguments");
    table_mbis_kernel<<<(nshell + 127) / 128, 128, 0, as_stream(stream)>>>(nshell, propars, shell_A,
                                                                           shell_alpha);
    HP_LAUNCH_CHECK("table_mbis_kernel");
    return HP_OK;
}

extern "C" int hp_table_scaled(int32_t nshell, const double* coeffs, const double* norms,
                               double* shell_A, void* stream) {
    HP_REQUIRE(nshell > 0 && coeffs && norms && shell_A, "bad arguments");
    table_scaled_kernel<<<(nshell + 127) / 128, 128, 0, as_stream(stream)>>>(nshell, coeffs, norms,
                                                                             shell_A);
    HP_LAUNCH_CHECK("table_scaled_kernel");
    return HP_OK;
}

extern "C" int hp_table_nlis(int32_t nshell, const double* propars, const double* inv_gamma,
                             double* shell_A, double* shell_alpha, double* shell_order,
                             void* stream) {
    HP_REQUIRE(nshell > 0 && propars && inv_gamma && shell_A && shell_alpha && shell_order,
               "bad arguments");
    table_nlis_kernel<<<(nshell + 127) / 128, 128, 0, as_stream(stream)>>>(
        nshell, propars, inv_gamma, shell_A, shell_alpha, shell_order);
    HP_LAUNCH_CHECK("table_nlis_kernel");
    return HP_OK;
}

extern "C" int hp_shell_project(int32_t nshell, const int64_t* shell_point_offsets,
                                const double* at_weights, const double* rho, const double* atgrid_w,
                                const double* shell_r, const double* shell_r2w, double* out_sph_avg,
                                void* stream) {
    HP_REQUIRE(nshell >= 0, "bad sizes");
    if (nshell == 0) return HP_OK;
    HP_REQUIRE(shell_point_offsets && at_weights && rho && atgrid_w && shell_r && shell_r2w &&
                   out_sph_avg, "null input");
    const int warps_per_block = 8;
    int64_t blocks = (int64_t(nshell) + warps_per_block - 1) / warps_per_block;
    const int64_t cap = int64_t(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    shell_project_kernel<<<int(blocks), warps_per_block * 32, 0, as_stream(stream)>>>(
        nshell, shell_point_offsets, at_weights, rho, atgrid_w, shell_r, shell_r2w, out_sph_avg);
    HP_LAUNCH_CHECK("shell_project_kernel");
    return HP_OK;
}

extern "C" int hp_shell_harmonics(int32_t nshell, int32_t lmax, const int64_t* shell_point_offsets,
                                  const int32_t* shell_atom, const double* px, const double* py,
                                  const double* pz, const double* atom_xyz, const double* at_weights,
                                  const double* rho, const double* atgrid_w, const double* shell_r,
                                  const double* shell_r2w, double* out, void* stream) {
    HP_REQUIRE(nshell >= 0, "bad sizes");
    HP_REQUIRE(lmax >= 0 && lmax <= kMaxHarmL, "lmax must be in 0..16");
    if (nshell == 0) return HP_OK;
    HP_REQUIRE(shell_point_offsets && shell_atom && px && py && pz && atom_xyz && at_weights && rho &&
                   atgrid_w && shell_r && shell_r2w && out, "null input");
    const int warps_per_block = 4;
    const int blocks = (nshell + warps_per_block - 1) / warps_per_block;
    shell_harmonics_kernel<<<blocks, warps_per_block * 32, 0, as_stream(stream)>>>(
        nshell, lmax, shell_point_offsets, shell_atom, px, py, pz, atom_xyz, at_weights, rho, atgrid_w,
        shell_r, shell_r2w, out);
    HP_LAUNCH_CHECK("shell_harmonics_kernel");
    return HP_OK;
}

extern "C" int hp_mbis_radial_solve(int32_t natom, int32_t atom_base, const int32_t* rad_offsets, const double* rad_r,
                                    const double* rad_w4, const double* sph_avg,
                                    const int32_t* par_offsets, double* propars,
                                    const double* pseudo_numbers, double inner_threshold,
                                    double density_cutoff, int32_t max_inner, int32_t nrad_max,
                                    int32_t nshell_max, double* charges, double* msd, int32_t* niter,
                                    uint32_t* flags, void* stream) {
    HP_REQUIRE(natom >= 0, "bad sizes");
    if (natom == 0) return HP_OK;
    HP_REQUIRE(rad_offsets && rad_r && rad_w4 && sph_avg && par_offsets && propars &&
                   pseudo_numbers && charges && msd && niter && flags, "null input");
    HP_REQUIRE(nrad_max > 0 && nrad_max <= 4096, "nrad_max must be in 1..4096");
    HP_REQUIRE(nshell_max > 0 && nshell_max <= kMaxMbisShells, "nshell_max must be in 1..7");
    // shared memory: one double per radial point of the largest atom; the caller guarantees
    // nrad <= 4096 (checked on the Python side where the offsets live on the host)
    static const bool force_warp = [] { const char* e = getenv("HP_B200_RADIAL_WARP"); return e && e[0] == '1'; }();
    if (!force_warp && nrad_max <= kRadThreads * kRadOwn) {
        // shells per atom: <= 3 up to argon (mbis.py:36-46); the small instantiation halves the code size
        if (nshell_max <= 3)
            shell_fixed_point_block_kernel<false, 3><<<natom, kRadThreads, 0, as_stream(stream)>>>(
                natom, atom_base, rad_offsets, rad_r, rad_w4, sph_avg, par_offsets, propars, nullptr, nullptr,
                pseudo_numbers, inner_threshold, density_cutoff, max_inner, charges, msd, niter, flags);
        else
            shell_fixed_point_block_kernel<false, kMaxMbisShells><<<natom, kRadThreads, 0, as_stream(stream)>>>(
                natom, atom_base, rad_offsets, rad_r, rad_w4, sph_avg, par_offsets, propars, nullptr, nullptr,
                pseudo_numbers, inner_threshold, density_cutoff, max_inner, charges, msd, niter, flags);
        HP_LAUNCH_CHECK("shell_fixed_point_block_kernel<mbis>");
        return HP_OK;
    }
    const size_t smem = sizeof(double) * 4096;
    mbis_radial_kernel<<<natom, 32, smem, as_stream(stream)>>>(
        natom, atom_base, rad_offsets, rad_r, rad_w4, sph_avg, par_offsets, propars, pseudo_numbers,
        inner_threshold, density_cutoff, max_inner, charges, msd, niter, flags);
    HP_LAUNCH_CHECK("mbis_radial_kernel");
    return HP_OK;
}

extern "C" int hp_finish_iteration(int32_t npartial, const double* entropy_partials, int32_t natom,
                                   const double* msd, double* out2, void* stream) {
    HP_REQUIRE(natom > 0 && msd && out2, "bad arguments");
    finish_iteration_kernel<<<1, 256, 0, as_stream(stream)>>>(npartial, entropy_partials, natom, msd,
                                                              out2);
    HP_LAUNCH_CHECK("finish_iteration_kernel");
    return HP_OK;
}

extern "C" int hp_nlis_radial_solve(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                                    const double* rad_r, const double* rad_w4, const double* sph_avg,
                                    const int32_t* par_offsets, double* propars,
                                    const int32_t* shell_offsets, const double* inv_gamma,
                                    const double* pseudo_numbers, double inner_threshold,
                                    double density_cutoff, int32_t max_inner, int32_t nrad_max,
                                    int32_t nshell_max, double* charges, double* msd, int32_t* niter,
                                    uint32_t* flags, void* stream) {
    HP_REQUIRE(natom >= 0, "bad sizes");
    if (natom == 0) return HP_OK;
    HP_REQUIRE(rad_offsets && rad_r && rad_w4 && sph_avg && par_offsets && propars &&
                   shell_offsets && inv_gamma && pseudo_numbers && charges && msd && niter && flags,
               "null input");
    HP_REQUIRE(nrad_max > 0 && nrad_max <= 4096, "nrad_max must be in 1..4096");
    HP_REQUIRE(nshell_max > 0 && nshell_max <= kMaxMbisShells, "nshell_max must be in 1..7");
    static const bool force_warp = [] { const char* e = getenv("HP_B200_RADIAL_WARP"); return e && e[0] == '1'; }();
    if (!force_warp && nrad_max <= kRadThreads * kRadOwn) {
        if (nshell_max <= 3)
            shell_fixed_point_block_kernel<true, 3><<<natom, kRadThreads, 0, as_stream(stream)>>>(
                natom, atom_base, rad_offsets, rad_r, rad_w4, sph_avg, par_offsets, propars, shell_offsets, inv_gamma,
                pseudo_numbers, inner_threshold, density_cutoff, max_inner, charges, msd, niter, flags);
        else
            shell_fixed_point_block_kernel<true, kMaxMbisShells><<<natom, kRadThreads, 0, as_stream(stream)>>>(
                natom, atom_base, rad_offsets, rad_r, rad_w4, sph_avg, par_offsets, propars, shell_offsets, inv_gamma,
                pseudo_numbers, inner_threshold, density_cutoff, max_inner, charges, msd, niter, flags);
        HP_LAUNCH_CHECK("shell_fixed_point_block_kernel<nlis>");
        return HP_OK;
    }
    nlis_radial_kernel<<<natom, 32, sizeof(double) * 4096, as_stream(stream)>>>(
        natom, atom_base, rad_offsets, rad_r, rad_w4, sph_avg, par_offsets, propars, shell_offsets,
        inv_gamma, pseudo_numbers, inner_threshold, density_cutoff, max_inner, charges, msd, niter,
        flags);
    HP_LAUNCH_CHECK("nlis_radial_kernel");
    return HP_OK;
}

extern "C" int hp_lisa_sc_radial_solve(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                                       const double* rad_w4, const double* sph_avg,
                                       const int32_t* par_offsets, double* propars,
                                       const int64_t* bs_offsets, const double* bs_funcs,
                                       const double* pseudo_numbers, double inner_threshold,
                                       double density_cutoff, double population_cutoff,
                                       int32_t max_inner, int32_t single_update, int32_t nrad_max,
                                       int32_t nshell_max, double* charges, double* msd,
                                       int32_t* niter, uint32_t* flags, void* stream) {
    HP_REQUIRE(natom >= 0, "bad sizes");
    if (natom == 0) return HP_OK;
    HP_REQUIRE(rad_offsets && rad_w4 && sph_avg && par_offsets && propars && bs_offsets && bs_funcs &&
                   pseudo_numbers && charges && msd && niter && flags, "null input");
    HP_REQUIRE(nrad_max > 0 && nshell_max > 0, "bad shared-memory sizes");
#define HP_SC_BLOCK(KPW, NPL)                                                                            \
    lisa_sc_block_kernel<KPW, NPL><<<natom, kScThreads, 0, as_stream(stream)>>>(                         \
        natom, atom_base, rad_offsets, rad_w4, sph_avg, par_offsets, propars, bs_offsets, bs_funcs,     \
        pseudo_numbers, inner_threshold, density_cutoff, population_cutoff, max_inner, single_update,   \
        charges, msd, niter, flags)
    // block-per-atom kernel for the usual shapes (K <= 24 shells, <= 256 radial points)
    // (HP_B200_SC_GENERIC=1 forces the one-warp-per-atom kernel: A/B runs and shape-fallback tests)
    static const bool force_generic = [] { const char* e = getenv("HP_B200_SC_GENERIC"); return e && e[0] == '1'; }();
    if (!force_generic && nshell_max <= 6 * kScWarps && nrad_max <= 256) {
        const bool small_k = nshell_max <= 4 * kScWarps;
        if (nrad_max <= 160) {
            if (small_k) HP_SC_BLOCK(4, 5); else HP_SC_BLOCK(6, 5);
        } else {
            if (small_k) HP_SC_BLOCK(4, 8); else HP_SC_BLOCK(6, 8);
        }
        HP_LAUNCH_CHECK("lisa_sc_block_kernel");
        return HP_OK;
    }
#undef HP_SC_BLOCK
    const size_t smem = sizeof(double) * (2 * size_t(nrad_max) + 2 * size_t(nshell_max));
    HP_REQUIRE(smem <= 200 * 1024, "radial grid / basis too large for shared memory");
    if (smem > 48 * 1024) {
        int rc = check_cuda(cudaFuncSetAttribute(lisa_sc_radial_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                            "cudaFuncSetAttribute");
        if (rc) return rc;
    }
    lisa_sc_radial_kernel<<<natom, 32, smem, as_stream(stream)>>>(
        natom, atom_base, rad_offsets, rad_w4, sph_avg, par_offsets, propars, bs_offsets, bs_funcs,
        pseudo_numbers, inner_threshold, density_cutoff, population_cutoff, max_inner, single_update,
        nrad_max, charges, msd, niter, flags);
    HP_LAUNCH_CHECK("lisa_sc_radial_kernel");
    return HP_OK;
}

extern "C" int hp_atom_moments(int32_t natom, int32_t atom_base, int32_t lmax,
                               const int64_t* seg_offsets, const double* px, const double* py,
                               const double* pz, const double* atgrid_w, const double* at_weights,
                               const double* dens, const double* atom_xyz, double* out, void* stream) {
    HP_REQUIRE(natom >= 0, "bad sizes");
    HP_REQUIRE(lmax >= 0 && lmax <= kMaxL, "lmax must be in 0..4");
    if (natom == 0) return HP_OK;
    HP_REQUIRE(seg_offsets && px && py && pz && atgrid_w && at_weights && dens && atom_xyz && out,
               "null input");
    const int nmom = (lmax + 1) * (lmax + 2) * (lmax + 3) / 6 + (lmax + 1) * (lmax + 1) + (lmax + 1);
    atom_moments_kernel<<<natom, 128, 0, as_stream(stream)>>>(natom, atom_base, lmax, seg_offsets, px, py,
                                                              pz, atgrid_w, at_weights, dens, atom_xyz,
                                                              nmom, out);
    HP_LAUNCH_CHECK("atom_moments_kernel");
    return HP_OK;
}
