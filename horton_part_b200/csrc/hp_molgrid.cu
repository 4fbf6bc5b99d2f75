// Molecular-grid reductions for the global schemes (row a12 of SURVEY.md section 8a: gLISA) and the
// molecular-grid variants of the per-atom updates (row a9, grid_type 2/3).
//
//   hp_shell_moments   I_m = sum_p t(p) * A_m * exp(-alpha_m r_pm^n),  t = molw*rho/rho0^power (masked)
//                      power 1: function_g (glisa.py:850-879: new c_m = c_m * I_m) and the gradient
//                      (glisa.py:454-458: grad_m = -I_m);  A_m = the shell's normalisation.
//   hp_atom_weight_integrals   N_a = sum_p molw*rho*clip(rho0_a/rho0, 0, 1)   (glisa.py:269-278)
//   hp_hessian         H_mn = sum_p u(p) g_m(p) g_n(p), u = molw*rho/rho0^2 (masked)  (glisa.py:459-470)
//
// Basis functions are regenerated from coordinates + parameters per tile: the reference's
// (M, Npts) `pro_shells` / `rho*pro_shells` arrays (glisa.py:335-344; 2 x 105 GB at config 4) are
// never materialised.  Partial sums are kept per thread block and combined in a fixed order, so
// results are bit-reproducible run to run.
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_math.cuh"

namespace hp {

constexpr int kMgThreads = 256;
constexpr int kMgWarps = kMgThreads / 32;
constexpr int kMgTileAtoms = 128;
constexpr int kMgTileShells = 1024;

struct __align__(16) MgAtom {
    double x, y, z;
    int s0, ns;
};

template <int F>
__device__ __forceinline__ double mg_radial(double d2) {
    return (F == HP_FUNCTOR_GAUSS) ? d2 : sqrt_nocall(d2);
}

template <int F>
__device__ __forceinline__ double mg_shell(double alpha, double n, double r) {
    if (F == HP_FUNCTOR_GENERAL) {
        const double rn = (n == 1.0) ? r : ((n == 2.0) ? r * r : pow(r, n));
        return exp(-alpha * rn);
    }
    return exp_neg_poly(-alpha * r);
}

// Screening of the moments pass (MODE 0).  For one chunk of points with bounding sphere (c, R) every point
// is at least dmin = max(|c - R_a| - R, 0) away from atom a, so shell k of that atom adds at most
// |A_k| exp(-alpha_k dmin^n) sum_p |t(p)| to out[k] over the whole chunk.  Below 2^-kMgScreenBits the
// (chunk, shell) pair is skipped by the whole block: the integrals I_k = int rho g_k / rho0 are O(1) (they are
// 1 at the gLISA fixed point) and a grid has < 2^20 chunks, so everything skipped together stays below
// 2^-60 -- under the rounding of the sums themselves.  HP_B200_MOMENTS_SCREEN=0 evaluates every pair.
constexpr double kMgScreenBits = 80.0;

template <int kP>
__device__ __forceinline__ void mg_chunk_bounds(const double (&x)[kP], const double (&y)[kP], const double (&z)[kP],
                                                const double (&t)[kP], int64_t first, int64_t npts,
                                                double (*s_box)[kMgWarps], double& cx, double& cy, double& cz,
                                                double& crad, double& log_tsum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double big = 1e300;
    double lo[3] = {big, big, big}, hi[3] = {-big, -big, -big}, ts = 0.0;
#pragma unroll
    for (int j = 0; j < kP; ++j) {
        if (first + int64_t(j) * kMgThreads + threadIdx.x < npts) {
            lo[0] = fmin(lo[0], x[j]); hi[0] = fmax(hi[0], x[j]);
            lo[1] = fmin(lo[1], y[j]); hi[1] = fmax(hi[1], y[j]);
            lo[2] = fmin(lo[2], z[j]); hi[2] = fmax(hi[2], z[j]);
        }
        ts += fabs(t[j]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
        }
        ts += __shfl_xor_sync(0xffffffffu, ts, off);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_box[a][warp] = lo[a]; s_box[3 + a][warp] = hi[a]; }
        s_box[6][warp] = ts;
    }
    __syncthreads();
    ts = 0.0;
    for (int w = 0; w < kMgWarps; ++w) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], s_box[a][w]); hi[a] = fmax(hi[a], s_box[3 + a][w]); }
        ts += s_box[6][w];
    }
    cx = 0.5 * (lo[0] + hi[0]); cy = 0.5 * (lo[1] + hi[1]); cz = 0.5 * (lo[2] + hi[2]);
    // radius: farthest point of the chunk from the box centre (tighter than the half diagonal)
    double r2 = 0.0;
#pragma unroll
    for (int j = 0; j < kP; ++j) {
        if (first + int64_t(j) * kMgThreads + threadIdx.x < npts) {
            const double dx = x[j] - cx, dy = y[j] - cy, dz = z[j] - cz;
            r2 = fmax(r2, fma(dz, dz, fma(dy, dy, dx * dx)));
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, off));
    if (lane == 0) s_box[7][warp] = r2;
    __syncthreads();
    for (int w = 0; w < kMgWarps; ++w) r2 = fmax(r2, s_box[7][w]);
    crad = sqrt(r2) * (1.0 + 1e-12);
    log_tsum = log(ts);  // -inf for an all-masked chunk: every shell is dead
}

// MODE 0: per-shell moments with weight t;  MODE 1: per-atom clipped-weight integrals.
template <int F, int MODE, int kP>
__global__ void __launch_bounds__(kMgThreads, 2)
molgrid_reduce_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                      const double* __restrict__ pz, int natom, const double* __restrict__ atom_xyz,
                      const int* __restrict__ atom_sh_off, const double* __restrict__ shell_A,
                      const double* __restrict__ shell_alpha, const double* __restrict__ shell_order,
                      int ntile, const int* __restrict__ tile_off, const double* __restrict__ rho,
                      const double* __restrict__ molw, const double* __restrict__ promol,
                      double density_cutoff, int power, int nout, int screen, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // MODE 0 screening (see mg_chunk_bounds): per chunk, shells that cannot contribute are skipped block-wide
    __shared__ double s_box[8][kMgWarps];
    __shared__ double s_dmin[kMgTileAtoms];
    __shared__ unsigned char s_live[kMgTileShells];
    __shared__ unsigned char s_alive[kMgTileAtoms];
    MgAtom* s_atoms = reinterpret_cast<MgAtom*>(smem_raw);
    double2* s_AB = reinterpret_cast<double2*>(s_atoms + kMgTileAtoms);
    double* s_N = reinterpret_cast<double*>(s_AB + kMgTileShells);
    double* s_acc = s_N + ((F == HP_FUNCTOR_GENERAL) ? kMgTileShells : 0);  // [kMgWarps][kMgTileShells]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* my_partial = partial + int64_t(blockIdx.x) * nout;
    for (int i = threadIdx.x; i < nout; i += kMgThreads) my_partial[i] = 0.0;

    const int64_t span = int64_t(kMgThreads) * kP;
    const int64_t nchunk = (npts + span - 1) / span;
    for (int64_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        double x[kP], y[kP], z[kP], t[kP], inv[kP];
#pragma unroll
        for (int j = 0; j < kP; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kMgThreads + threadIdx.x;
            const bool live = p < npts;
            const int64_t q = live ? p : npts - 1;
            x[j] = px[q]; y[j] = py[q]; z[j] = pz[q];
            const double r0 = promol[q], rh = rho[q];
            if (MODE == 0) {
                const bool sick = (rh < density_cutoff) || (r0 < density_cutoff);
                double v = sick ? 0.0 : molw[q] * rh / r0;
                if (power == 2 && !sick) v /= r0;
                t[j] = live ? v : 0.0;
                inv[j] = 0.0;
            } else {
                t[j] = live ? molw[q] * rh : 0.0;
                inv[j] = r0;
            }
        }
        // bounding sphere of the chunk's points and sum |t| over them (MODE 0 with screening)
        double cx = 0.0, cy = 0.0, cz = 0.0, crad = 0.0, log_tsum = 0.0;
        const bool screened = (MODE == 0) && screen;
        if (screened) mg_chunk_bounds<kP>(x, y, z, t, chunk * span, npts, s_box, cx, cy, cz, crad, log_tsum);
        for (int tl = 0; tl < ntile; ++tl) {
            const int a0 = tile_off[tl], a1 = tile_off[tl + 1];
            const int sh0 = atom_sh_off[a0], sh1 = atom_sh_off[a1];
            __syncthreads();
            for (int i = threadIdx.x; i < a1 - a0; i += kMgThreads) {
                MgAtom rec;
                rec.x = atom_xyz[3 * (a0 + i)]; rec.y = atom_xyz[3 * (a0 + i) + 1]; rec.z = atom_xyz[3 * (a0 + i) + 2];
                rec.s0 = atom_sh_off[a0 + i] - sh0;
                rec.ns = atom_sh_off[a0 + i + 1] - atom_sh_off[a0 + i];
                s_atoms[i] = rec;
                if (screened) {
                    const double dx = cx - rec.x, dy = cy - rec.y, dz = cz - rec.z;
                    s_dmin[i] = fmax(sqrt(fma(dz, dz, fma(dy, dy, dx * dx))) - crad, 0.0);
                }
            }
            for (int i = threadIdx.x; i < sh1 - sh0; i += kMgThreads) {
                s_AB[i] = make_double2(shell_A[sh0 + i], shell_alpha[sh0 + i]);
                if (F == HP_FUNCTOR_GENERAL) s_N[i] = shell_order[sh0 + i];
            }
            __syncthreads();
            if (screened) {
                // shell k of atom i is dead for this chunk when |A_k| exp(-alpha_k dmin^n) sum|t| < 2^-kMgScreenBits
                for (int i = threadIdx.x; i < a1 - a0; i += kMgThreads) {
                    const MgAtom rec = s_atoms[i];
                    const double d = s_dmin[i];
                    unsigned char any = 0;
                    for (int k = 0; k < rec.ns; ++k) {
                        const double2 ab = s_AB[rec.s0 + k];
                        const double n = (F == HP_FUNCTOR_GENERAL) ? s_N[rec.s0 + k] : 1.0;
                        const double dn = (F == HP_FUNCTOR_GAUSS) ? d * d
                                          : ((F == HP_FUNCTOR_SLATER || n == 1.0) ? d : ((n == 2.0) ? d * d : pow(d, n)));
                        // log(|A| sum|t|) - alpha d^n >= -bits ln 2  (a NaN anywhere keeps the shell)
                        const bool dead = log(fabs(ab.x)) + log_tsum - ab.y * dn < -kMgScreenBits * 0.6931471805599453;
                        s_live[rec.s0 + k] = dead ? 0 : 1;
                        any |= dead ? 0 : 1;
                    }
                    s_alive[i] = any;
                }
                __syncthreads();
            }
            for (int i = 0; i < a1 - a0; ++i) {
                if (screened && !s_alive[i]) continue;
                const MgAtom rec = s_atoms[i];
                double r[kP], f[kP];
#pragma unroll
                for (int j = 0; j < kP; ++j) {
                    const double dx = x[j] - rec.x, dy = y[j] - rec.y, dz = z[j] - rec.z;
                    r[j] = mg_radial<F>(fma(dz, dz, fma(dy, dy, dx * dx)));
                    f[j] = 0.0;
                }
                for (int k = 0; k < rec.ns; ++k) {
                    if (screened && !s_live[rec.s0 + k]) continue;
                    const double2 ab = s_AB[rec.s0 + k];
                    const double n = (F == HP_FUNCTOR_GENERAL) ? s_N[rec.s0 + k] : 1.0;
                    if (MODE == 0) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < kP; ++j) s = fma(t[j], mg_shell<F>(ab.y, n, r[j]), s);
                        s = warp_allsum(s * ab.x);
                        if (lane == 0) s_acc[warp * kMgTileShells + rec.s0 + k] = s;
                    } else {
#pragma unroll
                        for (int j = 0; j < kP; ++j) f[j] = fma(ab.x, mg_shell<F>(ab.y, n, r[j]), f[j]);
                    }
                }
                if (MODE == 1) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < kP; ++j) s = fma(t[j], fmin(fmax(f[j] / inv[j], 0.0), 1.0), s);
                    s = warp_allsum(s);
                    if (lane == 0) s_acc[warp * kMgTileShells + i] = s;
                }
            }
            __syncthreads();
            const int ncol = (MODE == 0) ? (sh1 - sh0) : (a1 - a0);
            const int col0 = (MODE == 0) ? sh0 : a0;
            for (int c = threadIdx.x; c < ncol; c += kMgThreads) {
                if (screened && !s_live[c]) continue;  // nothing was accumulated for a dead shell
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kMgWarps; ++w) tot += s_acc[w * kMgTileShells + c];
                my_partial[col0 + c] += tot;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// grid_type 2/3 (row a9 on the molecular grid): one inner iteration of the per-atom fixed points
// for ALL atoms at once.  Three shell-parameter sets over the same (alpha-order) structure:
//   out  : outer parameters -> w_a = clip(rho0_a/promol, 0, 1), rhoa = w_a*rho  (fixed during the
//          inner loop; mbis.py:170-173, gisa.py:257-260)
//   in   : current inner parameters -> terms t_k = A_k exp(-alpha_k r^n), pro = sum_k t_k
//   prev : previous inner parameters -> oldpro (the reference keeps the array, we recompute it)
// Per shell:  S0_k = sum molw t_k ratio,  S1_k = sum molw t_k ratio r^n   (ratio = rhoa/pro, masked)
// Per atom :  chg = sum molw (oldpro - pro)^2,  pop = sum molw rhoa
// (mbis.py:128-152, alisa.py:262-274 with weights = grid.weights, r = radial_distances[a]).
// Atoms with active[a] == 0 are skipped.  With in = new and prev = old outer parameters the chg
// column is the atom's term of compute_change on the molecular grid (core/iterstock.py:40-41).
// ---------------------------------------------------------------------------------------------
constexpr int kUpTileAtoms = 32;
constexpr int kUpTileShells = 128;
constexpr int kUpCols = 2 * kUpTileShells + 2 * kUpTileAtoms;

struct __align__(16) UpAtom {
    double x, y, z;
    int s0, ns;
};

template <int F>
__global__ void __launch_bounds__(kMgThreads, 2)
molgrid_update_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                      const double* __restrict__ pz, int natom, const double* __restrict__ atom_xyz,
                      const int* __restrict__ atom_sh_off, const double* __restrict__ A_out,
                      const double* __restrict__ al_out, const double* __restrict__ A_in,
                      const double* __restrict__ al_in, const double* __restrict__ A_prev,
                      const double* __restrict__ al_prev, const double* __restrict__ shell_order,
                      const int* __restrict__ active, int ntile, const int* __restrict__ tile_off,
                      const double* __restrict__ rho, const double* __restrict__ molw,
                      const double* __restrict__ promol, double density_cutoff, int nshell,
                      double* __restrict__ partial) {
    constexpr int kP = 2;
    __shared__ UpAtom s_atoms[kUpTileAtoms];
    __shared__ int s_active[kUpTileAtoms];
    __shared__ double s_par[6][kUpTileShells];  // A_out al_out A_in al_in A_prev al_prev
    __shared__ double s_N[kUpTileShells];
    __shared__ double s_acc[kMgWarps][kUpCols];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nout = 2 * nshell + 2 * natom;
    double* my_partial = partial + int64_t(blockIdx.x) * nout;
    for (int i = threadIdx.x; i < nout; i += kMgThreads) my_partial[i] = 0.0;

    const int64_t span = int64_t(kMgThreads) * kP;
    const int64_t nchunk = (npts + span - 1) / span;
    for (int64_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        double x[kP], y[kP], z[kP], rh[kP], mw[kP], pm[kP];
#pragma unroll
        for (int j = 0; j < kP; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kMgThreads + threadIdx.x;
            const bool live = p < npts;
            const int64_t q = live ? p : npts - 1;
            x[j] = px[q]; y[j] = py[q]; z[j] = pz[q];
            rh[j] = rho[q]; pm[j] = promol[q];
            mw[j] = live ? molw[q] : 0.0;
        }
        for (int tl = 0; tl < ntile; ++tl) {
            const int a0 = tile_off[tl], a1 = tile_off[tl + 1];
            const int sh0 = atom_sh_off[a0], sh1 = atom_sh_off[a1];
            __syncthreads();
            for (int i = threadIdx.x; i < a1 - a0; i += kMgThreads) {
                UpAtom rec;
                rec.x = atom_xyz[3 * (a0 + i)]; rec.y = atom_xyz[3 * (a0 + i) + 1]; rec.z = atom_xyz[3 * (a0 + i) + 2];
                rec.s0 = atom_sh_off[a0 + i] - sh0;
                rec.ns = atom_sh_off[a0 + i + 1] - atom_sh_off[a0 + i];
                s_atoms[i] = rec;
                s_active[i] = active ? active[a0 + i] : 1;
            }
            for (int i = threadIdx.x; i < sh1 - sh0; i += kMgThreads) {
                s_par[0][i] = A_out[sh0 + i]; s_par[1][i] = al_out[sh0 + i];
                s_par[2][i] = A_in[sh0 + i];  s_par[3][i] = al_in[sh0 + i];
                s_par[4][i] = A_prev[sh0 + i]; s_par[5][i] = al_prev[sh0 + i];
                s_N[i] = (F == HP_FUNCTOR_GENERAL) ? shell_order[sh0 + i] : 1.0;
            }
            for (int i = threadIdx.x; i < kUpCols; i += kMgThreads)
                for (int w = 0; w < kMgWarps; ++w) s_acc[w][i] = 0.0;
            __syncthreads();
            for (int i = 0; i < a1 - a0; ++i) {
                if (!s_active[i]) continue;
                const UpAtom rec = s_atoms[i];
                double r[kP], rhoa[kP], pro[kP], old[kP];
#pragma unroll
                for (int j = 0; j < kP; ++j) {
                    const double dx = x[j] - rec.x, dy = y[j] - rec.y, dz = z[j] - rec.z;
                    r[j] = mg_radial<F>(fma(dz, dz, fma(dy, dy, dx * dx)));
                    double yo = 0.0;
                    pro[j] = old[j] = 0.0;
                    for (int k = 0; k < rec.ns; ++k) {
                        const int c = rec.s0 + k;
                        yo = fma(s_par[0][c], mg_shell<F>(s_par[1][c], s_N[c], r[j]), yo);
                        pro[j] += s_par[2][c] * mg_shell<F>(s_par[3][c], s_N[c], r[j]);
                        old[j] += s_par[4][c] * mg_shell<F>(s_par[5][c], s_N[c], r[j]);
                    }
                    rhoa[j] = fmin(fmax(yo / pm[j], 0.0), 1.0) * rh[j];
                }
                double chg = 0.0, pop = 0.0;
#pragma unroll
                for (int j = 0; j < kP; ++j) {
                    const double e = old[j] - pro[j];
                    chg = fma(mw[j] * e, e, chg);
                    pop = fma(mw[j], rhoa[j], pop);
                }
                chg = warp_allsum(chg);
                pop = warp_allsum(pop);
                if (lane == 0) {
                    s_acc[warp][2 * kUpTileShells + 2 * i] = chg;
                    s_acc[warp][2 * kUpTileShells + 2 * i + 1] = pop;
                }
                for (int k = 0; k < rec.ns; ++k) {
                    const int c = rec.s0 + k;
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int j = 0; j < kP; ++j) {
                        const bool sick = (rhoa[j] < density_cutoff) || (pro[j] < density_cutoff);
                        const double ratio = sick ? 0.0 : rhoa[j] / pro[j];
                        const double t = s_par[2][c] * mg_shell<F>(s_par[3][c], s_N[c], r[j]) * ratio;
                        const double rn = (F == HP_FUNCTOR_GAUSS) ? r[j]
                                          : ((s_N[c] == 1.0) ? r[j] : pow(r[j], s_N[c]));
                        s0 = fma(mw[j], t, s0);
                        s1 = fma(mw[j] * t, rn, s1);
                    }
                    s0 = warp_allsum(s0);
                    s1 = warp_allsum(s1);
                    if (lane == 0) {
                        s_acc[warp][2 * c] = s0;
                        s_acc[warp][2 * c + 1] = s1;
                    }
                }
            }
            __syncthreads();
            for (int c = threadIdx.x; c < 2 * (sh1 - sh0); c += kMgThreads) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kMgWarps; ++w) tot += s_acc[w][c];
                my_partial[2 * sh0 + c] += tot;
            }
            for (int c = threadIdx.x; c < 2 * (a1 - a0); c += kMgThreads) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kMgWarps; ++w) tot += s_acc[w][2 * kUpTileShells + c];
                my_partial[2 * nshell + 2 * a0 + c] += tot;
            }
        }
    }
}

// out[c] = sum over rows of partial[row][c], rows added in order.
__global__ void __launch_bounds__(256)
reduce_rows_kernel(int nrows, int ncols, const double* __restrict__ partial, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double s = 0.0;
    for (int r = 0; r < nrows; ++r) s += partial[int64_t(r) * ncols + c];
    out[c] = s;
}

// msd_a = int 4 pi r^2 (rho0_a[c_new] - rho0_a[c_old])^2 on atom a's radial grid, basis functions
// tabulated on the grid (K x nrad); one warp per atom.  compute_change for the exponential-basis
// schemes (core/iterstock.py:32-45 with gisa.py:42-65).
__global__ void __launch_bounds__(32)
radial_change_kernel(int natom, int atom_base, const int* __restrict__ rad_off,
                     const double* __restrict__ rad_w4, const int* __restrict__ par_off,
                     const int64_t* __restrict__ bs_off, const double* __restrict__ bs,
                     const double* __restrict__ c_new, const double* __restrict__ c_old,
                     double* __restrict__ msd) {
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x, lane = threadIdx.x;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    const int p0 = par_off[a], K = par_off[a + 1] - p0;
    const double* g = bs + bs_off[blockIdx.x];
    double dev = 0.0;
    for (int i = lane; i < nrad; i += 32) {
        double yn = 0.0, yo = 0.0;
        for (int k = 0; k < K; ++k) {
            yn += c_new[p0 + k] * g[k * nrad + i];
            yo += c_old[p0 + k] * g[k * nrad + i];
        }
        const double d = yn - yo;
        dev += rad_w4[r0 + i] * d * d;
    }
    dev = warp_allsum(dev);
    if (lane == 0) msd[a] = dev;
}

// Line-search admissibility of candidate coefficient vectors (glisa.py:283-307 is_promol_valid /
// is_proatom_valid on the radial grids): for candidate j and atom a
//   flag = 1 if any rho0_a(r_i) < negative_cutoff, else
//   flag = 2 if check_mono and any rho0_a(r_i) - rho0_a(r_{i+1}) < negative_cutoff, else 0.
// grid = (local atoms, candidates), one warp each; candidates are rows of `cand` (ncand x npar).
__global__ void __launch_bounds__(32)
radial_valid_kernel(int natom, int atom_base, const int* __restrict__ rad_off,
                    const int* __restrict__ par_off, const int64_t* __restrict__ bs_off,
                    const double* __restrict__ bs, const double* __restrict__ cand, int npar,
                    double negative_cutoff, int check_mono, int* __restrict__ flags) {
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x, lane = threadIdx.x, j = blockIdx.y;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    const int p0 = par_off[a], K = par_off[a + 1] - p0;
    const double* g = bs + bs_off[blockIdx.x];
    const double* c = cand + int64_t(j) * npar + p0;
    bool negative = false, rising = false;
    // lane handles i = lane, lane+32, ...; the value at i+1 is recomputed (K is small)
    for (int i = lane; i < nrad; i += 32) {
        double y = 0.0, ynext = 0.0;
        const bool has_next = i + 1 < nrad;
        for (int k = 0; k < K; ++k) {
            y += c[k] * g[k * nrad + i];
            if (has_next) ynext += c[k] * g[k * nrad + i + 1];
        }
        negative |= y < negative_cutoff;
        if (has_next) rising |= (y - ynext) < negative_cutoff;
    }
    negative = __any_sync(0xffffffffu, negative);
    rising = __any_sync(0xffffffffu, rising);
    if (lane == 0) flags[int64_t(j) * natom + blockIdx.x] = negative ? 1 : ((check_mono && rising) ? 2 : 0);
}

template <int F, int MODE, int kP>
static int launch_molgrid(int64_t npts, const double* px, const double* py, const double* pz,
                          int natom, const double* atom_xyz, const int* atom_sh_off,
                          const double* shell_A, const double* shell_alpha, const double* shell_order,
                          int ntile, const int* tile_off, const double* rho, const double* molw,
                          const double* promol, double cutoff, int power, int nout, int nblocks,
                          double* partial, cudaStream_t st) {
    const size_t smem = sizeof(MgAtom) * kMgTileAtoms + sizeof(double2) * kMgTileShells +
                        sizeof(double) * ((F == HP_FUNCTOR_GENERAL) ? kMgTileShells : 0) +
                        sizeof(double) * kMgWarps * kMgTileShells;
    auto kern = molgrid_reduce_kernel<F, MODE, kP>;
    int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                        "cudaFuncSetAttribute");
    if (rc) return rc;
    const char* env = getenv("HP_B200_MOMENTS_SCREEN");  // read per call: tests compare both settings
    const int screen = !(env && env[0] == '0');
    kern<<<nblocks, kMgThreads, smem, st>>>(npts, px, py, pz, natom, atom_xyz, atom_sh_off, shell_A,
                                            shell_alpha, shell_order, ntile, tile_off, rho, molw, promol,
                                            cutoff, power, nout, screen, partial);
    return check_cuda(cudaGetLastError(), "molgrid_reduce_kernel");
}

}  // namespace hp

using namespace hp;

extern "C" int32_t hp_molgrid_num_blocks(int64_t npts) {
    const int64_t span = int64_t(kMgThreads) * 4;
    int64_t want = (npts + span - 1) / span;
    const int64_t cap = int64_t(sm_count()) * 2;
    if (want > cap) want = cap;
    return int32_t(want < 1 ? 1 : want);
}

static int molgrid_dispatch(int mode, int functor, int64_t npts, const double* px, const double* py,
                            const double* pz, int natom, const double* atom_xyz, const int* atom_sh_off,
                            const double* shell_A, const double* shell_alpha, const double* shell_order,
                            int ntile, const int* tile_off, const double* rho, const double* molw,
                            const double* promol, double cutoff, int power, int nout, double* partial,
                            double* out, void* stream) {
    cudaStream_t st = as_stream(stream);
    const int nblocks = hp_molgrid_num_blocks(npts);
    int rc = HP_ERR_ARG;
#define HP_MG(F, MODE, KP)                                                                              \
    rc = launch_molgrid<F, MODE, KP>(npts, px, py, pz, natom, atom_xyz, atom_sh_off, shell_A, shell_alpha, \
                                     shell_order, ntile, tile_off, rho, molw, promol, cutoff, power, nout, \
                                     nblocks, partial, st)
    if (mode == 0) {
        if (functor == HP_FUNCTOR_SLATER) HP_MG(HP_FUNCTOR_SLATER, 0, 4);
        else if (functor == HP_FUNCTOR_GAUSS) HP_MG(HP_FUNCTOR_GAUSS, 0, 4);
        else if (functor == HP_FUNCTOR_GENERAL) HP_MG(HP_FUNCTOR_GENERAL, 0, 4);
    } else {
        if (functor == HP_FUNCTOR_SLATER) HP_MG(HP_FUNCTOR_SLATER, 1, 4);
        else if (functor == HP_FUNCTOR_GAUSS) HP_MG(HP_FUNCTOR_GAUSS, 1, 4);
        else if (functor == HP_FUNCTOR_GENERAL) HP_MG(HP_FUNCTOR_GENERAL, 1, 4);
    }
#undef HP_MG
    if (rc == HP_ERR_ARG) set_error("molgrid reduction: unsupported functor %d", functor);
    if (rc) return rc;
    reduce_rows_kernel<<<(nout + 255) / 256, 256, 0, st>>>(nblocks, nout, partial, out);
    return check_cuda(cudaGetLastError(), "reduce_rows_kernel");
}

extern "C" int hp_shell_moments(int functor, int64_t npts, const double* px, const double* py,
                                const double* pz, int32_t natom, const double* atom_xyz,
                                const int32_t* atom_shell_offsets, const double* shell_A,
                                const double* shell_alpha, const double* shell_order, int32_t ntile,
                                const int32_t* tile_atom_offsets, const double* rho, const double* molw,
                                const double* promol, double density_cutoff, int32_t power,
                                int32_t nshell, double* partial, double* out, void* stream) {
    HP_REQUIRE(npts > 0 && natom > 0 && ntile > 0 && nshell > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_shell_offsets && shell_A && shell_alpha &&
                   tile_atom_offsets && rho && molw && promol && partial && out, "null input");
    HP_REQUIRE(power == 1 || power == 2, "power must be 1 or 2");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    return molgrid_dispatch(0, functor, npts, px, py, pz, natom, atom_xyz, atom_shell_offsets, shell_A,
                            shell_alpha, shell_order, ntile, tile_atom_offsets, rho, molw, promol,
                            density_cutoff, power, nshell, partial, out, stream);
}

extern "C" int hp_atom_weight_integrals(int functor, int64_t npts, const double* px, const double* py,
                                        const double* pz, int32_t natom, const double* atom_xyz,
                                        const int32_t* atom_shell_offsets, const double* shell_A,
                                        const double* shell_alpha, const double* shell_order,
                                        int32_t ntile, const int32_t* tile_atom_offsets,
                                        const double* rho, const double* molw, const double* promol,
                                        double* partial, double* out, void* stream) {
    HP_REQUIRE(npts > 0 && natom > 0 && ntile > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_shell_offsets && shell_A && shell_alpha &&
                   tile_atom_offsets && rho && molw && promol && partial && out, "null input");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    return molgrid_dispatch(1, functor, npts, px, py, pz, natom, atom_xyz, atom_shell_offsets, shell_A,
                            shell_alpha, shell_order, ntile, tile_atom_offsets, rho, molw, promol, 0.0, 1,
                            natom, partial, out, stream);
}

extern "C" int hp_radial_change(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                                const double* rad_w4, const int32_t* par_offsets,
                                const int64_t* bs_offsets, const double* bs_funcs, const double* c_new,
                                const double* c_old, double* msd, void* stream) {
    HP_REQUIRE(natom >= 0, "bad sizes");
    if (natom == 0) return HP_OK;
    HP_REQUIRE(rad_offsets && rad_w4 && par_offsets && bs_offsets && bs_funcs && c_new && c_old && msd,
               "null input");
    radial_change_kernel<<<natom, 32, 0, as_stream(stream)>>>(natom, atom_base, rad_offsets, rad_w4,
                                                              par_offsets, bs_offsets, bs_funcs, c_new,
                                                              c_old, msd);
    HP_LAUNCH_CHECK("radial_change_kernel");
    return HP_OK;
}

extern "C" int hp_radial_valid(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                               const int32_t* par_offsets, const int64_t* bs_offsets,
                               const double* bs_funcs, const double* candidates, int32_t ncand,
                               int32_t npar, double negative_cutoff, int32_t check_mono, int32_t* flags,
                               void* stream) {
    HP_REQUIRE(natom >= 0 && ncand >= 0 && npar > 0, "bad sizes");
    if (natom == 0 || ncand == 0) return HP_OK;
    HP_REQUIRE(ncand <= 65535, "too many candidates");
    HP_REQUIRE(rad_offsets && par_offsets && bs_offsets && bs_funcs && candidates && flags, "null input");
    radial_valid_kernel<<<dim3(natom, ncand), 32, 0, as_stream(stream)>>>(
        natom, atom_base, rad_offsets, par_offsets, bs_offsets, bs_funcs, candidates, npar, negative_cutoff,
        check_mono, flags);
    HP_LAUNCH_CHECK("radial_valid_kernel");
    return HP_OK;
}

extern "C" void hp_molgrid_update_tile_limits(int32_t* max_atoms_host, int32_t* max_shells_host) {
    *max_atoms_host = kUpTileAtoms;
    *max_shells_host = kUpTileShells;
}

extern "C" int hp_molgrid_update_pass(int functor, int64_t npts, const double* px, const double* py,
                                      const double* pz, int32_t natom, const double* atom_xyz,
                                      const int32_t* atom_shell_offsets, const double* A_out,
                                      const double* alpha_out, const double* A_in, const double* alpha_in,
                                      const double* A_prev, const double* alpha_prev,
                                      const double* shell_order, const int32_t* active, int32_t ntile,
                                      const int32_t* tile_atom_offsets, const double* rho,
                                      const double* molw, const double* promol, double density_cutoff,
                                      int32_t nshell, double* partial, double* out, void* stream) {
    HP_REQUIRE(npts > 0 && natom > 0 && ntile > 0 && nshell > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_shell_offsets && A_out && alpha_out && A_in && alpha_in &&
                   A_prev && alpha_prev && tile_atom_offsets && rho && molw && promol && partial && out,
               "null input");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    cudaStream_t st = as_stream(stream);
    const int nblocks = hp_molgrid_num_blocks(npts);
    const int nout = 2 * nshell + 2 * natom;
#define HP_UP(F)                                                                                        \
    molgrid_update_kernel<F><<<nblocks, kMgThreads, 0, st>>>(                                           \
        npts, px, py, pz, natom, atom_xyz, atom_shell_offsets, A_out, alpha_out, A_in, alpha_in, A_prev, \
        alpha_prev, shell_order, active, ntile, tile_atom_offsets, rho, molw, promol, density_cutoff,    \
        nshell, partial)
    switch (functor) {
        case HP_FUNCTOR_SLATER: HP_UP(HP_FUNCTOR_SLATER); break;
        case HP_FUNCTOR_GAUSS: HP_UP(HP_FUNCTOR_GAUSS); break;
        case HP_FUNCTOR_GENERAL: HP_UP(HP_FUNCTOR_GENERAL); break;
        default: set_error("hp_molgrid_update_pass: unsupported functor %d", functor); return HP_ERR_ARG;
    }
#undef HP_UP
    HP_LAUNCH_CHECK("molgrid_update_kernel");
    reduce_rows_kernel<<<(nout + 255) / 256, 256, 0, st>>>(nblocks, nout, partial, out);
    return check_cuda(cudaGetLastError(), "reduce_rows_kernel");
}
