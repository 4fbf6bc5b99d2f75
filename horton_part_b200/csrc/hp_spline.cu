// Spline pro-atoms: ISA, Hirshfeld, Hirshfeld-I (row a5 of SURVEY.md section 8a).
//
//   hp_spline_build          not-a-knot cubic spline per atom (SciPy CubicSpline semantics,
//                            core/stockholder.py:259-269), one thread per atom (Thomas solve)
//   hp_promol_weights_spline fused promolecule / owner-weight / entropy pass with piecewise-cubic
//                            pro-atoms evaluated at |r_p - R_a| (core/stockholder.py:271-350),
//                            extrapolating with the end pieces exactly like PPoly
//   hp_isa_update            ISA's parameter update: propars = clipped spherical average, charge,
//                            change term (isa.py:102-122, core/iterstock.py:32-45)
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "hp_common.cuh"
#include "hp_math.cuh"

namespace hp {

// ---------------------------------------------------------------------------------------------
// Spline construction.  For knots x_0..x_{n-1} and values y, solve for the knot derivatives s with
// the not-a-knot end conditions (the banded system SciPy assembles), then emit PPoly coefficients
// (c0, c1, c2, c3) per interval:  S(x) = c0 d^3 + c1 d^2 + c2 d + c3,  d = x - x_i.
// `work` provides 2 doubles of scratch per knot.
// ---------------------------------------------------------------------------------------------
__global__ void spline_build_kernel(int natom, const int* __restrict__ knot_off,
                                    const double* __restrict__ knots, const double* __restrict__ values,
                                    int clip_negative, double* __restrict__ coef,
                                    double* __restrict__ work) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= natom) return;
    const int o = knot_off[a], n = knot_off[a + 1] - o;
    const double* x = knots + o;
    const double* yv = values + o;
    double* cp = work + 2 * o;       // modified upper diagonal
    double* s = work + 2 * o + n;    // rhs -> solution (knot derivatives)
    double* c = coef + 4 * (o - a);  // atom a owns n-1 segments; segments of all atoms are packed
    auto Y = [&](int i) {
        const double v = yv[i];
        return (clip_negative && v < 0.0) ? 0.0 : v;  // fix_proatom_rho, core/stockholder.py:218-219
    };
    if (n < 2) return;
    if (n == 2) {  // straight line
        const double slope = (Y(1) - Y(0)) / (x[1] - x[0]);
        c[0] = 0.0; c[1] = 0.0; c[2] = slope; c[3] = Y(0);
        return;
    }
    if (n == 3) {  // parabola through three points (SciPy's special case)
        const double dx0 = x[1] - x[0], dx1 = x[2] - x[1];
        const double sl0 = (Y(1) - Y(0)) / dx0, sl1 = (Y(2) - Y(1)) / dx1;
        // A = [[1,1,0],[dx1, 2(dx0+dx1), dx0],[0,1,1]], b = [2 sl0, 3(dx0 sl1 + dx1 sl0), 2 sl1]
        const double b0 = 2 * sl0, b1 = 3 * (dx0 * sl1 + dx1 * sl0), b2 = 2 * sl1;
        // eliminate: s0 = b0 - s1, s2 = b2 - s1
        const double s1 = (b1 - dx1 * b0 - dx0 * b2) / (2 * (dx0 + dx1) - dx1 - dx0);
        s[0] = b0 - s1; s[1] = s1; s[2] = b2 - s1;
    } else {
        // row 0:  dx1 * s0 + (x2 - x0) * s1 = b0
        const double dx0 = x[1] - x[0], dx1 = x[2] - x[1];
        const double sl0 = (Y(1) - Y(0)) / dx0, sl1 = (Y(2) - Y(1)) / dx1;
        double d = x[2] - x[0];
        double diag = dx1, upper = d;
        double rhs = ((dx0 + 2 * d) * dx1 * sl0 + dx0 * dx0 * sl1) / d;
        cp[0] = upper / diag;
        s[0] = rhs / diag;
        // interior rows i:  dx_i * s_{i-1} + 2 (dx_{i-1} + dx_i) * s_i + dx_{i-1} * s_{i+1} = b_i
        double dxm = dx0, slm = sl0;  // dx_{i-1}, slope_{i-1}
        for (int i = 1; i < n - 1; ++i) {
            const double dxi = x[i + 1] - x[i];
            const double sli = (Y(i + 1) - Y(i)) / dxi;
            const double lower = dxi;
            diag = 2 * (dxm + dxi) - lower * cp[i - 1];
            rhs = 3 * (dxi * slm + dxm * sli) - lower * s[i - 1];
            cp[i] = dxm / diag;
            s[i] = rhs / diag;
            dxm = dxi;
            slm = sli;
        }
        // last row:  (x_{n-1} - x_{n-3}) * s_{n-2} + dx_{n-3} * s_{n-1} = b_{n-1}
        const double dxl = x[n - 1] - x[n - 2], dxl2 = x[n - 2] - x[n - 3];
        const double sll = (Y(n - 1) - Y(n - 2)) / dxl, sll2 = (Y(n - 2) - Y(n - 3)) / dxl2;
        d = x[n - 1] - x[n - 3];
        const double lower = d;
        diag = dxl2 - lower * cp[n - 2];
        rhs = (dxl * dxl * sll2 + (2 * d + dxl) * dxl2 * sll) / d - lower * s[n - 2];
        s[n - 1] = rhs / diag;
        for (int i = n - 2; i >= 0; --i) s[i] -= cp[i] * s[i + 1];
    }
    for (int i = 0; i < n - 1; ++i) {
        const double dxi = x[i + 1] - x[i];
        const double slope = (Y(i + 1) - Y(i)) / dxi;
        const double t = (s[i] + s[i + 1] - 2 * slope) / dxi;
        c[4 * i + 0] = t / dxi;
        c[4 * i + 1] = (slope - s[i]) / dxi - t;
        c[4 * i + 2] = s[i];
        c[4 * i + 3] = Y(i);
    }
}

// Parallel version for the iteration loop: the not-a-knot system matrix depends on the KNOTS only, which
// never change during a partitioning -- only the right-hand side (the tabulated values) does.  The host
// inverts the matrix once per distinct knot array (hp_spline_system_inverse, stored transposed so that the
// matrix-vector product below reads it coalesced); one block per atom then builds the right-hand side, the
// knot derivatives s = A^-1 b and the PPoly coefficients with all threads.  The serial Thomas solve above
// (one thread per atom, ~300 dependent FP64 operations through global scratch) was 62 % of an ISA
// iteration at config 2 (profiles/r2_launches_*.txt).  inv_offsets[a] < 0 selects the serial path for
// that atom (fewer than 4 knots).
constexpr int kSbThreads = 256;

__global__ void __launch_bounds__(kSbThreads)
spline_build_inv_kernel(int natom, const int* __restrict__ knot_off, const double* __restrict__ knots,
                        const double* __restrict__ values, int clip_negative,
                        const long long* __restrict__ inv_off, const double* __restrict__ invT,
                        double* __restrict__ coef) {
    extern __shared__ double s_sb[];  // y[n] | rhs[n] | s[n]
    const int a = blockIdx.x;
    if (a >= natom) return;
    const int o = knot_off[a], n = knot_off[a + 1] - o;
    const double* x = knots + o;
    double* y = s_sb;
    double* rhs = s_sb + n;
    double* sd = s_sb + 2 * n;
    double* c = coef + 4 * (long long)(o - a);
    for (int i = threadIdx.x; i < n; i += kSbThreads) {
        const double v = values[o + i];
        y[i] = (clip_negative && v < 0.0) ? 0.0 : v;  // fix_proatom_rho, core/stockholder.py:218-219
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kSbThreads) {
        double b;
        if (i == 0) {
            const double dx0 = x[1] - x[0], dx1 = x[2] - x[1], d = x[2] - x[0];
            const double sl0 = (y[1] - y[0]) / dx0, sl1 = (y[2] - y[1]) / dx1;
            b = ((dx0 + 2 * d) * dx1 * sl0 + dx0 * dx0 * sl1) / d;
        } else if (i == n - 1) {
            const double dxl = x[n - 1] - x[n - 2], dxl2 = x[n - 2] - x[n - 3], d = x[n - 1] - x[n - 3];
            const double sll = (y[n - 1] - y[n - 2]) / dxl, sll2 = (y[n - 2] - y[n - 3]) / dxl2;
            b = (dxl * dxl * sll2 + (2 * d + dxl) * dxl2 * sll) / d;
        } else {
            const double dxm = x[i] - x[i - 1], dxi = x[i + 1] - x[i];
            const double slm = (y[i] - y[i - 1]) / dxm, sli = (y[i + 1] - y[i]) / dxi;
            b = 3 * (dxi * slm + dxm * sli);
        }
        rhs[i] = b;
    }
    __syncthreads();
    const double* AT = invT + inv_off[a];  // AT[j * n + i] = (A^-1)[i][j]
    for (int i = threadIdx.x; i < n; i += kSbThreads) {
        // eight partial sums, sixteen independent loads per trip: the product is bound by the latency of the
        // (L2-resident) matrix reads, not by arithmetic
        double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        int j = 0;
        for (; j + 15 < n; j += 16) {
            double m[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) m[u] = AT[(long long)(j + u) * n + i];
#pragma unroll
            for (int u = 0; u < 16; ++u) acc[u & 7] = fma(m[u], rhs[j + u], acc[u & 7]);
        }
        for (; j < n; ++j) acc[j & 7] = fma(AT[(long long)j * n + i], rhs[j], acc[j & 7]);
        const double s0 = acc[0] + acc[4], s1 = acc[1] + acc[5], s2 = acc[2] + acc[6], s3 = acc[3] + acc[7];
        sd[i] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n - 1; i += kSbThreads) {
        const double dxi = x[i + 1] - x[i];
        const double slope = (y[i + 1] - y[i]) / dxi;
        const double t = (sd[i] + sd[i + 1] - 2 * slope) / dxi;
        c[4 * i + 0] = t / dxi;
        c[4 * i + 1] = (slope - sd[i]) / dxi - t;
        c[4 * i + 2] = sd[i];
        c[4 * i + 3] = y[i];
    }
}

// PPoly evaluation with extrapolation (interval = clamp(searchsorted_right(x, r) - 1, 0, n-2)).
__device__ __forceinline__ double spline_eval(const double* __restrict__ x, const double* __restrict__ c,
                                              int n, double r) {
    int lo = 0, hi = n - 1;  // invariant: answer in [lo, hi-1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] <= r) lo = mid; else hi = mid;
    }
    const double d = r - x[lo];
    const double* ci = c + 4 * lo;
    double z = d;
    double res = fma(ci[2], z, ci[3]);
    z *= d;
    res = fma(ci[1], z, res);
    z *= d;
    return fma(ci[0], z, res);
}

// ---------------------------------------------------------------------------------------------
// Fused promolecule / owner-weight / entropy pass for piecewise-cubic pro-atoms.
//
// Interval index without a binary search: the top bits of the double r (sign, exponent and the first
// kLutMantBits mantissa bits) are a piecewise-linear log2 r, monotone in r, so
//     bin = clamp((hi32(r) >> kLutShift) - key0, 0, nbins - 1),   i = lut[bin]
// is the interval that contains the lower edge of r's bin; a forward scan `while x[i+1] <= r` (0-1
// steps for the radial transforms of qc-grid at 32 bins per octave, bounded by the knot count for
// any grid) then lands on searchsorted_right(x, r) - 1 clamped to [0, n-2], i.e. exactly PPoly's
// interval incl. extrapolation with the end pieces.  The table is built on the host once per distinct
// knot array (hp_spline_lut_size / hp_spline_lut_fill); atoms with the same radial grid share it.
//
// Atoms are streamed through shared memory in tiles (knots + PPoly coefficients, <= kSplTileKnots
// knots per tile); a thread owns kSplPts points of a chunk of consecutive points and keeps their
// running promolecule sums in registers, summing in atom order like the reference.
// ---------------------------------------------------------------------------------------------
constexpr int kSplThreads = 256;
constexpr int kSplTileKnots = 1536;  // 1536 knots + 4 x 1536 coefficients = 61,440 B of dynamic shared memory: 3 blocks per SM
constexpr int kSplTileAtoms = 64;
constexpr int kLutMantBits = 5;
constexpr int kLutShift = 20 - kLutMantBits;

struct SplAtom {  // 48 bytes
    double x, y, z;
    int ko, n;       // first knot in the tile's knot array, number of knots
    int key0, nbins;
    long long lut;   // offset of the atom's table in the LUT pool
};

template <int PTS>
__global__ void __launch_bounds__(kSplThreads, 3)
promol_weights_spline_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                             const double* __restrict__ pz, int64_t point_base, int natom,
                             const double* __restrict__ atom_xyz, const int64_t* __restrict__ atom_pt_off,
                             const int* __restrict__ knot_off, const double* __restrict__ knots,
                             const double* __restrict__ coef, const int* __restrict__ lut_meta,
                             const unsigned short* __restrict__ lut, int ntile,
                             const int* __restrict__ tile_off, double proatom_offset, double promol_offset,
                             const double* __restrict__ rho, const double* __restrict__ molw,
                             double density_cutoff, double* __restrict__ promol_out,
                             double* __restrict__ w_out, double* __restrict__ entropy_partials,
                             int npartial) {
    extern __shared__ __align__(16) double s_dyn[];  // coefficients (4 per segment) | knots
    double* s_coef = s_dyn;
    double* s_knots = s_dyn + 4 * kSplTileKnots;
    __shared__ SplAtom s_atoms[kSplTileAtoms];
    __shared__ double s_red[32];
    __shared__ int s_own[2];
    const int64_t span = int64_t(kSplThreads) * PTS;
    const int64_t nchunk = (npts + span - 1) / span;
    double entropy_acc = 0.0;
    for (int64_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        // owner atoms of the chunk's first and last point (block-uniform); a per-point search is only
        // needed when the chunk straddles atom blocks
        if (threadIdx.x < 2) {
            const int64_t last = (chunk + 1) * span - 1;
            const int64_t g = point_base + (threadIdx.x == 0 ? chunk * span : (last < npts ? last : npts - 1));
            int lo = 0, hi = natom;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (atom_pt_off[mid] <= g) lo = mid; else hi = mid;
            }
            s_own[threadIdx.x] = lo;
        }
        __syncthreads();
        const int own_lo = s_own[0], own_hi = s_own[1];
        double x[PTS], y[PTS], z[PTS], pro[PTS], own[PTS];
        int owner[PTS];
#pragma unroll
        for (int j = 0; j < PTS; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kSplThreads + threadIdx.x;
            const int64_t q = p < npts ? p : npts - 1;
            x[j] = px[q]; y[j] = py[q]; z[j] = pz[q];
            pro[j] = 0.0; own[j] = 0.0;
            int lo = own_lo;
            if (own_hi != own_lo) {
                const int64_t g = point_base + q;
                int hi = own_hi + 1;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (atom_pt_off[mid] <= g) lo = mid; else hi = mid;
                }
            }
            owner[j] = lo;
        }
        for (int t = 0; t < ntile; ++t) {
            const int a0 = tile_off[t], a1 = tile_off[t + 1];
            const int k0 = knot_off[a0], k1 = knot_off[a1];
            __syncthreads();  // previous tile consumed
            // knots of atoms a0..a1-1 are contiguous, and so are their coefficient blocks
            for (int i = threadIdx.x; i < k1 - k0; i += kSplThreads) s_knots[i] = knots[k0 + i];
            {
                const double2* src = reinterpret_cast<const double2*>(coef + 4 * int64_t(k0 - a0));
                double2* dst = reinterpret_cast<double2*>(s_coef);
                const int n2 = 2 * ((k1 - k0) - (a1 - a0));
                for (int i = threadIdx.x; i < n2; i += kSplThreads) dst[i] = src[i];
            }
            if (threadIdx.x < a1 - a0) {
                const int a = a0 + threadIdx.x;
                SplAtom rec;
                rec.x = atom_xyz[3 * a]; rec.y = atom_xyz[3 * a + 1]; rec.z = atom_xyz[3 * a + 2];
                rec.ko = knot_off[a] - k0;
                rec.n = knot_off[a + 1] - knot_off[a];
                rec.key0 = lut_meta[3 * a];
                rec.nbins = lut_meta[3 * a + 1];
                rec.lut = lut_meta[3 * a + 2];
                s_atoms[threadIdx.x] = rec;
            }
            __syncthreads();
            for (int ia = 0; ia < a1 - a0; ++ia) {
                const SplAtom at = s_atoms[ia];
                const int a = a0 + ia;
                const double* xk = s_knots + at.ko;
                const double* ck = s_coef + 4 * (at.ko - ia);
                const unsigned short* tab = lut + at.lut;
                const int last = at.n - 2;
                // the PTS points of a thread go through every stage together (distance -> table look-up ->
                // scan -> knot / coefficient loads -> Horner), so that their dependent shared-memory and L1
                // latencies overlap instead of adding up
                double r[PTS];
                int idx[PTS];
#pragma unroll
                for (int j = 0; j < PTS; ++j) {
                    const double dx = x[j] - at.x, dy = y[j] - at.y, dz = z[j] - at.z;
                    r[j] = sqrt_nocall(fma(dz, dz, fma(dy, dy, dx * dx)));
                    int bin = (__double2hiint(r[j]) >> kLutShift) - at.key0;
                    bin = max(0, min(bin, at.nbins - 1));
                    idx[j] = tab[bin];
                }
                // forward scan: two branch-free steps cover the radial transforms of qc-grid at 32 bins per
                // octave (at most one knot per bin); the loop behind them is for arbitrary knot sets
                bool more = false;
#pragma unroll
                for (int j = 0; j < PTS; ++j) {
                    int i = idx[j];
                    i += (i < last && xk[i + 1] <= r[j]) ? 1 : 0;
                    i += (i < last && xk[i + 1] <= r[j]) ? 1 : 0;
                    more = more || (i < last && xk[i + 1] <= r[j]);
                    idx[j] = i;
                }
                if (__builtin_expect(more, 0)) {
#pragma unroll
                    for (int j = 0; j < PTS; ++j) {
                        int i = idx[j];
                        while (i < last && xk[i + 1] <= r[j]) ++i;
                        idx[j] = i;
                    }
                }
#pragma unroll
                for (int j = 0; j < PTS; ++j) {
                    const int i = idx[j];
                    const double d = r[j] - xk[i];
                    const double2 c01 = *reinterpret_cast<const double2*>(ck + 4 * i);
                    const double2 c23 = *reinterpret_cast<const double2*>(ck + 4 * i + 2);
                    // PPoly's evaluation order: c3 + c2 d + c1 d^2 + c0 d^3
                    double zz = d;
                    double res = fma(c23.x, zz, c23.y);
                    zz *= d;
                    res = fma(c01.y, zz, res);
                    zz *= d;
                    res = fma(c01.x, zz, res);
                    // eval_proatom (core/stockholder.py:343-349): spline(r) + 1e-100 ...
                    const double f = res + proatom_offset;
                    // ... update_pro (:169-170): promoldens += work; promoldens += 1e-100
                    pro[j] = (pro[j] + f) + promol_offset;
                    own[j] = (a == owner[j]) ? f : own[j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PTS; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kSplThreads + threadIdx.x;
            if (p >= npts) continue;
            if (promol_out) promol_out[p] = pro[j];
            if (w_out) w_out[p] = fmin(fmax(own[j] / pro[j], 0.0), 1.0);
            if (entropy_partials) {
                const double r = rho[p];
                const bool sick = (pro[j] < density_cutoff) || (r < density_cutoff);
                if (!sick) entropy_acc += molw[p] * r * log(r / pro[j]);
            }
        }
    }
    if (entropy_partials) {
        const double total = block_sum(entropy_acc, s_red);
        if (threadIdx.x == 0) entropy_partials[blockIdx.x] = total;
        if (blockIdx.x == 0)
            for (int i = gridDim.x + threadIdx.x; i < npartial; i += kSplThreads) entropy_partials[i] = 0.0;
    }
}

// N_a = sum_p molw[p] * dens[p] * clip((S_a(|r_p - R_a|) + offset) / promol[p], 0, 1) over ALL points of
// the slab: populations on the molecular grid for spline pro-atoms (grid_type 2/3; core/base.py:
// 287-298 with on_molgrid weights) without storing natom x Npts weight arrays.  blockIdx.y = atom,
// blockIdx.x strides over the points; the per-block sums are folded per atom in a fixed order.
constexpr int kAwiThreads = 256;

__global__ void __launch_bounds__(kAwiThreads)
atom_weight_integrals_spline_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                                    const double* __restrict__ pz, const double* __restrict__ atom_xyz,
                                    const int* __restrict__ knot_off, const double* __restrict__ knots,
                                    const double* __restrict__ coef, double proatom_offset,
                                    const double* __restrict__ dens, const double* __restrict__ molw,
                                    const double* __restrict__ promol, double* __restrict__ partial) {
    __shared__ double s_red[32];
    const int a = blockIdx.y;
    const double ax = atom_xyz[3 * a], ay = atom_xyz[3 * a + 1], az = atom_xyz[3 * a + 2];
    const int o = knot_off[a], n = knot_off[a + 1] - o;
    const double* xk = knots + o;
    const double* ck = coef + 4 * (o - a);
    double acc = 0.0;
    for (int64_t p = int64_t(blockIdx.x) * kAwiThreads + threadIdx.x; p < npts; p += int64_t(gridDim.x) * kAwiThreads) {
        const double dx = px[p] - ax, dy = py[p] - ay, dz = pz[p] - az;
        const double r = sqrt_nocall(fma(dz, dz, fma(dy, dy, dx * dx)));
        const double f = spline_eval(xk, ck, n, r) + proatom_offset;
        const double w = fmin(fmax(f / promol[p], 0.0), 1.0);
        acc += molw[p] * (w * dens[p]);
    }
    const double total = block_sum(acc, s_red);
    if (threadIdx.x == 0) partial[int64_t(a) * gridDim.x + blockIdx.x] = total;
}

__global__ void fold_atom_partials_kernel(int natom, int nblk, const double* __restrict__ partial,
                                          double* __restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= natom) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[int64_t(a) * nblk + b];
    out[a] = s;
}

// ISA update, one warp per atom (isa.py:102-122).
__global__ void __launch_bounds__(32)
isa_update_kernel(int natom, int atom_base, const int* __restrict__ rad_off,
                  const double* __restrict__ rad_r, const double* __restrict__ rad_w,
                  const double* __restrict__ sph, const int* __restrict__ par_off,
                  double* __restrict__ propars, const double* __restrict__ pseudo,
                  double* __restrict__ charges, double* __restrict__ msd) {
    if (int(blockIdx.x) >= natom) return;
    const int a = atom_base + blockIdx.x;
    const int lane = threadIdx.x;
    const int r0 = rad_off[blockIdx.x], nrad = rad_off[blockIdx.x + 1] - r0;
    double* par = propars + par_off[a];
    double pop = 0.0, dev = 0.0;
    for (int i = lane; i < nrad; i += 32) {
        const double r = fmin(fmax(rad_r[r0 + i], 1e-100), 1e10);   // isa.py:108
        const double v = fmax(sph[r0 + i], 1e-100);                 // isa.py:110
        const double shell = kFourPi * (r * r);                     // 4 pi r^2
        pop += rad_w[r0 + i] * (shell * v);                         // isa.py:120
        const double d = v - par[i];
        dev += rad_w[r0 + i] * shell * d * d;                       // core/iterstock.py:44
        par[i] = v;                                                 // isa.py:117
    }
    pop = warp_allsum(pop);
    dev = warp_allsum(dev);
    if (lane == 0) {
        charges[a] = pseudo[a] - pop;
        msd[a] = dev;
    }
}

}  // namespace hp

using namespace hp;

// Host helper: transposed inverse of the not-a-knot system of a knot array (n >= 4), n x n doubles:
// out[j * n + i] = (A^-1)[i][j], A as assembled by SciPy's CubicSpline(bc_type="not-a-knot") and by
// spline_build_kernel.  Gauss-Jordan with partial pivoting in long double, rounded to double at the end.
extern "C" int hp_spline_system_inverse(int32_t n, const double* knots_host, double* invT_host) {
    HP_REQUIRE(n >= 4 && knots_host && invT_host, "needs at least 4 knots");
    const double* x = knots_host;
    std::vector<long double> A(size_t(n) * n, 0.0L), B(size_t(n) * n, 0.0L);
    auto at = [&](std::vector<long double>& M, int r, int c) -> long double& { return M[size_t(r) * n + c]; };
    at(A, 0, 0) = (long double)x[2] - x[1];
    at(A, 0, 1) = (long double)x[2] - x[0];
    for (int i = 1; i < n - 1; ++i) {
        const long double dxm = (long double)x[i] - x[i - 1], dxi = (long double)x[i + 1] - x[i];
        at(A, i, i - 1) = dxi;
        at(A, i, i) = 2 * (dxm + dxi);
        at(A, i, i + 1) = dxm;
    }
    at(A, n - 1, n - 2) = (long double)x[n - 1] - x[n - 3];
    at(A, n - 1, n - 1) = (long double)x[n - 2] - x[n - 3];
    for (int i = 0; i < n; ++i) at(B, i, i) = 1.0L;
    for (int col = 0; col < n; ++col) {
        int piv = col;
        for (int r = col + 1; r < n; ++r)
            if (fabsl(at(A, r, col)) > fabsl(at(A, piv, col))) piv = r;
        HP_REQUIRE(at(A, piv, col) != 0.0L, "singular spline system (repeated knots?)");
        if (piv != col)
            for (int k = 0; k < n; ++k) {
                std::swap(at(A, piv, k), at(A, col, k));
                std::swap(at(B, piv, k), at(B, col, k));
            }
        const long double inv = 1.0L / at(A, col, col);
        for (int k = 0; k < n; ++k) {
            at(A, col, k) *= inv;
            at(B, col, k) *= inv;
        }
        for (int r = 0; r < n; ++r) {
            if (r == col) continue;
            const long double f = at(A, r, col);
            if (f == 0.0L) continue;
            for (int k = 0; k < n; ++k) {
                at(A, r, k) -= f * at(A, col, k);
                at(B, r, k) -= f * at(B, col, k);
            }
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) invT_host[size_t(j) * n + i] = double(at(B, i, j));
    return HP_OK;
}

extern "C" int hp_spline_build(int32_t natom, const int32_t* knot_offsets, const double* knots,
                               const double* values, int32_t clip_negative, double* coef,
                               double* work, const int64_t* inv_offsets, const double* invT,
                               int32_t nknot_max, void* stream) {
    HP_REQUIRE(natom > 0 && knot_offsets && knots && values && coef && work, "bad arguments");
    if (inv_offsets && invT && nknot_max >= 4 && nknot_max <= 2048) {
        // every atom of the launch has >= 4 knots (the caller passes inv_offsets only then)
        spline_build_inv_kernel<<<natom, kSbThreads, sizeof(double) * 3 * size_t(nknot_max), as_stream(stream)>>>(
            natom, knot_offsets, knots, values, clip_negative, reinterpret_cast<const long long*>(inv_offsets), invT,
            coef);
        HP_LAUNCH_CHECK("spline_build_inv_kernel");
        return HP_OK;
    }
    spline_build_kernel<<<(natom + 63) / 64, 64, 0, as_stream(stream)>>>(natom, knot_offsets, knots,
                                                                         values, clip_negative, coef, work);
    HP_LAUNCH_CHECK("spline_build_kernel");
    return HP_OK;
}

// ---- interval look-up tables (host) -------------------------------------------------------------
namespace {
inline int lut_key(double r) {
    long long bits;
    memcpy(&bits, &r, sizeof(bits));
    return int(bits >> (32 + hp::kLutShift));
}
// first key of the table and its number of bins for a knot array (capped: the scan covers the rest)
inline void lut_range(const double* x, int n, int* key0, int* nbins) {
    const int hi = lut_key(x[n - 1] > 0.0 ? x[n - 1] : 0.0);
    int lo = lut_key(x[0] > 0.0 ? x[0] : 0.0);
    const int cap = 4096;
    if (hi - lo + 1 > cap) lo = hi - cap + 1;
    *key0 = lo;
    *nbins = hi - lo + 1;
}
}  // namespace

extern "C" int32_t hp_spline_lut_size(int32_t nknot, const double* knots_host) {
    if (nknot < 2 || !knots_host) return 0;
    int key0, nbins;
    lut_range(knots_host, nknot, &key0, &nbins);
    return nbins;
}

extern "C" int hp_spline_lut_fill(int32_t nknot, const double* knots_host, int32_t* key0_out,
                                  uint16_t* lut_host) {
    HP_REQUIRE(nknot >= 2 && nknot <= 65535 && knots_host && key0_out && lut_host, "bad arguments");
    int key0, nbins;
    lut_range(knots_host, nknot, &key0, &nbins);
    *key0_out = key0;
    int i = 0;
    for (int b = 0; b < nbins; ++b) {
        // lower edge of bin b: the double whose top bits are key0 + b and whose other bits are 0
        const long long bits = (long long)(key0 + b) << (32 + hp::kLutShift);
        double edge;
        memcpy(&edge, &bits, sizeof(edge));
        while (i < nknot - 2 && knots_host[i + 1] <= edge) ++i;  // searchsorted_right(x, edge) - 1, clamped
        lut_host[b] = uint16_t(i);
    }
    // everything below the first bin is clamped into it: start the scan at the first interval there
    // (matters only when the number of bins was capped, i.e. key0 raised above the first knot's key)
    lut_host[0] = 0;
    return HP_OK;
}

extern "C" void hp_spline_tile_limits(int32_t* max_atoms_host, int32_t* max_knots_host) {
    *max_atoms_host = kSplTileAtoms;
    *max_knots_host = kSplTileKnots;
}

extern "C" int hp_promol_weights_spline(int64_t npts, const double* px, const double* py,
                                        const double* pz, int64_t point_base, int32_t natom,
                                        const double* atom_xyz, const int64_t* atom_point_offsets,
                                        const int32_t* knot_offsets, const double* knots,
                                        const double* coef, const int32_t* lut_meta, const uint16_t* lut,
                                        int32_t ntile, const int32_t* tile_atom_offsets,
                                        double proatom_offset, double promol_offset, const double* rho,
                                        const double* molw, double density_cutoff, double* promol,
                                        double* at_weights, double* entropy_partials, void* stream) {
    HP_REQUIRE(npts >= 0 && natom > 0 && ntile > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_point_offsets && knot_offsets && knots && coef,
               "null input");
    HP_REQUIRE(lut_meta && lut && tile_atom_offsets, "null look-up table / tiling");
    HP_REQUIRE(!entropy_partials || (rho && molw), "entropy needs rho and molw");
    const int npartial = hp_num_partials();
    if (npts == 0) {
        if (entropy_partials)
            return check_cuda(cudaMemsetAsync(entropy_partials, 0, sizeof(double) * npartial,
                                              as_stream(stream)), "memset partials");
        return HP_OK;
    }
    const size_t smem = sizeof(double) * 5 * kSplTileKnots;
    static bool configured[64] = {};  // per device
    if (first_use_on_device(configured)) {
        int rc = check_cuda(cudaFuncSetAttribute(promol_weights_spline_kernel<4>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)), "cudaFuncSetAttribute");
        if (rc == HP_OK)
            rc = check_cuda(cudaFuncSetAttribute(promol_weights_spline_kernel<2>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)), "cudaFuncSetAttribute");
        if (rc == HP_OK)
            rc = check_cuda(cudaFuncSetAttribute(promol_weights_spline_kernel<1>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)), "cudaFuncSetAttribute");
        if (rc) return rc;
    }
    // points per thread: 4 when that still gives >= 4 chunks per resident block, else 2 or 1 so that small
    // grids (config 2: 582,000 points = 569 chunks of 1,024 on 444 resident blocks) fill whole waves
    int64_t cap = int64_t(sm_count()) * 3;
    if (cap > npartial) cap = npartial;
    auto chunks = [&](int pts) { return (npts + int64_t(kSplThreads) * pts - 1) / (int64_t(kSplThreads) * pts); };
    const int pts = chunks(4) >= 4 * cap ? 4 : (chunks(2) >= 2 * cap ? 2 : 1);
    int64_t grid = chunks(pts);
    if (grid > cap) grid = cap;
#define HP_SPL_ARGS                                                                                        \
    npts, px, py, pz, point_base, natom, atom_xyz, atom_point_offsets, knot_offsets, knots, coef, lut_meta, \
        lut, ntile, tile_atom_offsets, proatom_offset, promol_offset, rho, molw, density_cutoff, promol,    \
        at_weights,                                                                                      \
        entropy_partials, npartial
    if (pts == 4) promol_weights_spline_kernel<4><<<int(grid), kSplThreads, smem, as_stream(stream)>>>(HP_SPL_ARGS);
    else if (pts == 2) promol_weights_spline_kernel<2><<<int(grid), kSplThreads, smem, as_stream(stream)>>>(HP_SPL_ARGS);
    else promol_weights_spline_kernel<1><<<int(grid), kSplThreads, smem, as_stream(stream)>>>(HP_SPL_ARGS);
#undef HP_SPL_ARGS
    HP_LAUNCH_CHECK("promol_weights_spline_kernel");
    return HP_OK;
}

extern "C" int hp_isa_update(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                             const double* rad_r, const double* rad_w, const double* sph_avg,
                             const int32_t* par_offsets, double* propars, const double* pseudo_numbers,
                             double* charges, double* msd, void* stream) {
    HP_REQUIRE(natom >= 0, "bad sizes");
    if (natom == 0) return HP_OK;
    HP_REQUIRE(rad_offsets && rad_r && rad_w && sph_avg && par_offsets && propars && pseudo_numbers &&
                   charges && msd, "null input");
    isa_update_kernel<<<natom, 32, 0, as_stream(stream)>>>(natom, atom_base, rad_offsets, rad_r, rad_w,
                                                           sph_avg, par_offsets, propars,
                                                           pseudo_numbers, charges, msd);
    HP_LAUNCH_CHECK("isa_update_kernel");
    return HP_OK;
}

extern "C" int32_t hp_spline_integral_blocks(int64_t npts) {
    const int64_t want = (npts + kAwiThreads - 1) / kAwiThreads;
    return int32_t(want < 1 ? 1 : (want > 64 ? 64 : want));
}

extern "C" int hp_atom_weight_integrals_spline(int64_t npts, const double* px, const double* py,
                                               const double* pz, int32_t natom, const double* atom_xyz,
                                               const int32_t* knot_offsets, const double* knots,
                                               const double* coef, double proatom_offset,
                                               const double* dens, const double* molw,
                                               const double* promol, double* partial, double* out,
                                               void* stream) {
    HP_REQUIRE(npts >= 0 && natom > 0 && atom_xyz && knot_offsets && knots && coef && partial && out,
               "bad arguments");
    HP_REQUIRE(npts == 0 || (px && py && pz && dens && molw && promol), "bad arguments");
    const int nblk = hp_spline_integral_blocks(npts);
    if (npts > 0) {
        atom_weight_integrals_spline_kernel<<<dim3(nblk, natom), kAwiThreads, 0, as_stream(stream)>>>(
            npts, px, py, pz, atom_xyz, knot_offsets, knots, coef, proatom_offset, dens, molw, promol, partial);
        HP_LAUNCH_CHECK("atom_weight_integrals_spline_kernel");
    } else {
        const int rc = check_cuda(cudaMemsetAsync(partial, 0, sizeof(double) * size_t(natom) * nblk,
                                                  as_stream(stream)), "memset partials");
        if (rc != HP_OK) return rc;
    }
    fold_atom_partials_kernel<<<(natom + 127) / 128, 128, 0, as_stream(stream)>>>(natom, nblk, partial, out);
    HP_LAUNCH_CHECK("fold_atom_partials_kernel");
    return HP_OK;
}
