// gLISA Hessian  H_mn = sum_p u(p) g_m(p) g_n(p),  u = molw*rho/rho0^2 masked  (glisa.py:459-470).
//
// The reference makes M(M+1)/2 NumPy passes over (M, Npts) tables.  Here the contraction
// B^T diag(u) B is done chunk by chunk over the grid points:
//   1. basis_chunk_kernel  regenerates  Gu[p][m] = sqrt(u_p) * g_m(p)  for one chunk of points into
//      a point-major scratch panel (row length Mpad = M rounded up to 128), so panel rows are
//      contiguous and the SYRK loads are fully coalesced;
//   2. syrk_panel_kernel   C_s += Gu_s^T Gu_s  on 128x128 tiles of the upper triangle, 8x8 register
//      micro-tiles of FP64 FMAs fed by a two-stage cp.async pipeline; the chunk is split over
//      `nsplit` sub-panels with one partial matrix each, nsplit chosen so that tiles x nsplit fills
//      whole waves of SMs, without atomics;
//   3. hessian_finish_kernel  H = sum_s C_s in a fixed order, mirrored to the lower triangle.
// Every element of every partial matrix is owned by exactly one block, so the result is
// bit-reproducible.  Bound: FP64 pipe, M(M+1) Npts flop (unit U2); B200 has no
// tcgen05 FP64 kind; the tile product runs on the FP64 tensor cores (mma.sync m8n8k4 = DMMA, see
// syrk_panel_dmma_kernel), the DFMA micro-kernel is kept behind HP_B200_HESSIAN_DFMA=1 for A/B runs.
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_math.cuh"
#include "hp_table.cuh"

namespace hp {

#ifndef HP_HESS_SUB
#define HP_HESS_SUB 1280  // points per sub-panel (measured at M = 1,500: 5120 -> 191 ms, 2560 -> 168 ms, 1280 -> 161 ms)
#endif
#ifndef HP_HESS_ROWS
#define HP_HESS_ROWS 32  // points per producer block
#endif
#ifndef HP_HESS_PBLOCKS
#define HP_HESS_PBLOCKS 3  // producer blocks per SM the register budget is set for
#endif
#ifndef HP_HESS_PANEL_MB
#define HP_HESS_PANEL_MB 1024  // bytes of one panel buffer (two of them)
#endif
constexpr int kHT = 128;   // tile edge
constexpr int kHQ = 64;    // screening granularity along the columns (a quadrant of a tile)
constexpr int kHK = 16;        // points per shared-memory slab
constexpr int kHMaxSplit = 64; // upper bound on sub-panels per chunk (partial matrices)

// Block screening, producer side: running maximum of |Gu| per (sub-panel, 64-column block) of the chunk, kept
// as the high word of the double (exponent + 20 mantissa bits: enough for a threshold test, and non-negative
// doubles order like their bit patterns).  Thread t of the 256 handles the columns t + 256 j, so the warp's
// j-th value lies in column block 4 j + warp / 2: one REDUX per value, lane j keeps the j-th result, and
// after the row the lanes publish their maxima with a guarded atomicMax (a plain L2 read first: almost
// every row is below the maximum already recorded).
struct PanelMax {
    unsigned keep = 0u;
    __device__ __forceinline__ void add(int j, double v) {
        const unsigned h = __reduce_max_sync(0xffffffffu, static_cast<unsigned>(__double2hiint(fabs(v))));
        if ((threadIdx.x & 31) == (j & 31)) keep = max(keep, h);
    }
    __device__ __forceinline__ void publish(unsigned long long* __restrict__ flags, int s, int nq, int nj) {
        const int j = threadIdx.x & 31;
        for (int jj = j; jj < nj; jj += 32) {  // nj <= 32 for Mpad <= 8192; beyond that the lanes hold merged maxima
            const int cb = 4 * jj + int(threadIdx.x >> 6);
            if (cb >= nq) continue;
            const unsigned long long bits = static_cast<unsigned long long>(keep) << 32;
            unsigned long long* f = flags + s * nq + cb;
            if (__ldcg(f) < bits) atomicMax(f, bits);
        }
    }
};

// Rows (points) per producer block: the shell parameters of a thread's columns are loaded once for all of
// them, and the block publishes its maxima once (one guarded atomic per 64-column block and kHRows rows).
constexpr int kHRows = HP_HESS_ROWS;
constexpr int kHRowUnroll = 8;  // rows in flight per thread (independent exponential chains)
static_assert(kHRows % kHRowUnroll == 0 && kHRows <= 256, "row groups");
static_assert(HP_HESS_SUB % kHRows == 0, "the rows of one producer block lie in one sub-panel");

// sqrt(u_p) and the coordinates of the block's rows -> shared memory (threads 0 .. kHRows-1)
__device__ __forceinline__ void panel_rows_setup(double (*s_pt)[4], int64_t p_first, int64_t npts,
                                                 const double* __restrict__ px, const double* __restrict__ py,
                                                 const double* __restrict__ pz, const double* __restrict__ rho,
                                                 const double* __restrict__ molw, const double* __restrict__ promol,
                                                 double cutoff) {
    if (threadIdx.x < kHRows) {
        const int64_t p = p_first + threadIdx.x;
        double su = 0.0, x = 0.0, y = 0.0, z = 0.0;
        if (p < npts) {
            const double r0 = promol[p], rh = rho[p];
            const bool sick = (rh < cutoff) || (r0 < cutoff);
            su = sick ? 0.0 : sqrt(molw[p] * rh / r0 / r0);
            x = px[p]; y = py[p]; z = pz[p];
        }
        s_pt[threadIdx.x][0] = su; s_pt[threadIdx.x][1] = x; s_pt[threadIdx.x][2] = y; s_pt[threadIdx.x][3] = z;
    }
    __syncthreads();
}

template <int F>
__global__ void __launch_bounds__(256, HP_HESS_PBLOCKS)
basis_chunk_kernel(int64_t p0, int pc, int M, int Mpad, const double* __restrict__ px,
                   const double* __restrict__ py, const double* __restrict__ pz,
                   const double* __restrict__ rho, const double* __restrict__ molw,
                   const double* __restrict__ promol, double cutoff,
                   const int* __restrict__ shell_atom, const double* __restrict__ atom_xyz,
                   const double* __restrict__ shell_norm, const double* __restrict__ shell_alpha,
                   const double* __restrict__ shell_order, int64_t npts, double* __restrict__ Gu,
                   unsigned long long* __restrict__ flags, int pc_sub) {
    // one block per kHRows points, threads over shells (coalesced stores along m)
    __shared__ double s_pt[kHRows][4];
    const int lp0 = blockIdx.x * kHRows;
    if (lp0 >= pc) return;
    panel_rows_setup(s_pt, p0 + lp0, npts, px, py, pz, rho, molw, promol, cutoff);
    double* rows = Gu + int64_t(lp0) * Mpad;
    PanelMax pmax;
    int j = 0;
    for (int m = threadIdx.x; m < Mpad; m += blockDim.x, ++j) {
        const bool valid = m < M;
        const int a = valid ? shell_atom[m] : 0;
        const double ax = atom_xyz[3 * a], ay = atom_xyz[3 * a + 1], az = atom_xyz[3 * a + 2];
        const double alpha = valid ? shell_alpha[m] : 0.0, norm = valid ? shell_norm[m] : 0.0;
        const double order = (valid && F == HP_FUNCTOR_GENERAL) ? shell_order[m] : 1.0;
        double vmax = 0.0;
        for (int r0 = 0; r0 < kHRows; r0 += kHRowUnroll)
#pragma unroll
        for (int rr = 0; rr < kHRowUnroll; ++rr) {
            const int r = r0 + rr;
            const double su = s_pt[r][0];
            double v = 0.0;
            if (valid && su != 0.0) {
                const double dx = s_pt[r][1] - ax, dy = s_pt[r][2] - ay, dz = s_pt[r][3] - az;
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                double e;
                if (F == HP_FUNCTOR_GAUSS) {
                    e = exp_neg_poly(-alpha * d2);
                } else if (F == HP_FUNCTOR_SLATER) {
                    e = exp_neg_poly(-alpha * sqrt_nocall(d2));
                } else {
                    const double dist = sqrt(d2);
                    const double rn = (order == 1.0) ? dist : ((order == 2.0) ? dist * dist : pow(dist, order));
                    e = exp(-alpha * rn);
                }
                v = su * norm * e;
            }
            rows[int64_t(r) * Mpad + m] = v;
            vmax = fmax(vmax, fabs(v));  // fmax drops a NaN operand: pass NaNs on explicitly
            if (v != v) vmax = __longlong_as_double(0x7ff0000000000000ll);
        }
        if (flags) pmax.add(j, vmax);
    }
    if (flags) pmax.publish(flags, lp0 / pc_sub, Mpad / kHQ, j);
}

// The same panel for tabulated basis functions (basis_type="numeric"): Gu[p][m] = sqrt(u_p) S_m(r_pm).
__global__ void __launch_bounds__(256)
table_basis_chunk_kernel(int64_t p0, int pc, int M, int Mpad, const double* __restrict__ px,
                         const double* __restrict__ py, const double* __restrict__ pz,
                         const double* __restrict__ rho, const double* __restrict__ molw,
                         const double* __restrict__ promol, double cutoff, const int* __restrict__ shell_atom,
                         const double* __restrict__ atom_xyz, TableArgs tab, int64_t npts,
                         double* __restrict__ Gu, unsigned long long* __restrict__ flags, int pc_sub) {
    __shared__ double s_pt[kHRows][4];
    const int lp0 = blockIdx.x * kHRows;
    if (lp0 >= pc) return;
    panel_rows_setup(s_pt, p0 + lp0, npts, px, py, pz, rho, molw, promol, cutoff);
    double* rows = Gu + int64_t(lp0) * Mpad;
    PanelMax pmax;
    int j = 0;
    for (int m = threadIdx.x; m < Mpad; m += blockDim.x, ++j) {
        const bool valid = m < M;
        const int a = valid ? shell_atom[m] : 0;
        const double ax = atom_xyz[3 * a], ay = atom_xyz[3 * a + 1], az = atom_xyz[3 * a + 2];
        const double* coef = tab.shell_coef + (valid ? tab.shell_coef_off[m] : 0);
        double vmax = 0.0;
#pragma unroll 2
        for (int r = 0; r < kHRows; ++r) {
            const double su = s_pt[r][0];
            double v = 0.0;
            if (valid && su != 0.0) {
                const double dx = s_pt[r][1] - ax, dy = s_pt[r][2] - ay, dz = s_pt[r][3] - az;
                double d;
                const int i = table_interval(tab, a, sqrt_nocall(fma(dz, dz, fma(dy, dy, dx * dx))), d);
                v = su * table_cubic(coef + 4 * i, d);
            }
            rows[int64_t(r) * Mpad + m] = v;
            vmax = fmax(vmax, fabs(v));
            if (v != v) vmax = __longlong_as_double(0x7ff0000000000000ll);
        }
        if (flags) pmax.add(j, vmax);
    }
    if (flags) pmax.publish(flags, lp0 / pc_sub, Mpad / kHQ, j);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// Block screening of the panel.  A chunk of points sees only the basis functions of the atoms around it:
// for most (sub-panel, 64-column block) pairs every entry of Gu is negligible.  The producer kernels record
// max |Gu| per (sub-panel s, column block q) while they write the panel (PanelMax above, no second pass over
// the panel), panel_chunk_max_kernel the chunk's overall maximum.  A 64 x 64 quadrant of a tile product is
// skipped when either of its column blocks stays below 2^-kHessScreenBits of that maximum: its contribution
// to any H_mn is then below 2^-kHessScreenBits of the largest term of the chunk, far below the rounding of
// the FP64 sums it would be added to (same argument as the atom screening of the promolecule kernel).  A
// block whose four quadrants are all skipped returns at once; in the DMMA kernel the warps of a skipped
// quadrant only take part in the operand pipeline (each warp scheduler holds one warp of every quadrant, so
// the tile's time follows the number of live quadrants), and the lower-left quadrant of a diagonal tile --
// never read by hessian_finish_kernel -- is always skipped.
// flags layout: [nsplit * nq] block maxima (nq = Mpad / 64) | [1] chunk maximum | [1] executed-quadrant
// counter (all uint64; non-negative doubles order like their bit patterns).  HP_B200_HESSIAN_SCREEN=0
// disables the skip.
// ---------------------------------------------------------------------------------------------
constexpr int kHessScreenBits = 64;

__global__ void __launch_bounds__(256)
panel_chunk_max_kernel(int nblock, unsigned long long* __restrict__ flags) {
    __shared__ unsigned long long s_red[8];
    unsigned long long m = 0ull;
    for (int i = threadIdx.x; i < nblock; i += blockDim.x) m = max(m, flags[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = max(m, s_red[w]);
        flags[nblock] = m;
    }
}

// Bit q = 2 i + j of the result: quadrant (row half i, column half j) of tile (tx, ty) of sub-panel s has
// to be computed.  Thread 0 adds the number of live quadrants to the counter.
__device__ __forceinline__ unsigned hessian_live_quadrants(const unsigned long long* __restrict__ flags, int nsplit,
                                                           int nq, int s, int2 tile) {
    unsigned live = (tile.x == tile.y) ? 0xbu : 0xfu;  // diagonal tile: rows 64.. x columns ..63 are never read
    if (!flags) return live;
    const unsigned long long gmax = flags[nsplit * nq];
    const unsigned long long drop = static_cast<unsigned long long>(kHessScreenBits) << 52;
    const unsigned long long thr = gmax > drop ? gmax - drop : 0ull;  // gmax * 2^-bits (0: keep everything)
    const unsigned long long* f = flags + s * nq;
    const bool r0 = f[2 * tile.x] >= thr, r1 = f[2 * tile.x + 1] >= thr;
    const bool c0 = f[2 * tile.y] >= thr, c1 = f[2 * tile.y + 1] >= thr;
    live &= (r0 && c0 ? 1u : 0u) | (r0 && c1 ? 2u : 0u) | (r1 && c0 ? 4u : 0u) | (r1 && c1 ? 8u : 0u);
    if (live && threadIdx.x == 0)
        atomicAdd(const_cast<unsigned long long*>(&flags[nsplit * nq + 1]), static_cast<unsigned long long>(__popc(live)));
    return live;
}

// C_s(tile) += Gu_s(:, tile.x)^T Gu_s(:, tile.y): 128x128 tile, 8x8 micro-tiles, the panel streamed
// through a two-stage cp.async pipeline of kHK-point slabs.
__global__ void __launch_bounds__(256)
syrk_panel_kernel(const double* __restrict__ Gu, int Mpad, int pc_sub, const int2* __restrict__ tiles,
                  double* __restrict__ Cpart, const unsigned long long* __restrict__ flags, int nq) {
    if (!hessian_live_quadrants(flags, gridDim.y, nq, blockIdx.y, tiles[blockIdx.x])) return;
    extern __shared__ __align__(16) double smem_syrk[];  // [2 stages][A|B][kHK][kHT]
    auto As = [&](int st, int kk) { return smem_syrk + ((st * 2 + 0) * kHK + kk) * kHT; };
    auto Bs = [&](int st, int kk) { return smem_syrk + ((st * 2 + 1) * kHK + kk) * kHT; };
    const int2 tile = tiles[blockIdx.x];
    const int s = blockIdx.y;
    const double* panel = Gu + int64_t(s) * pc_sub * Mpad;
    double* C = Cpart + int64_t(s) * Mpad * Mpad;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    // each thread moves (kHK*kHT*2 doubles)/256 threads in 16-byte pieces per slab
    auto issue = [&](int st, int k0) {
#pragma unroll
        for (int j = 0; j < (kHK * kHT) / (2 * 256); ++j) {
            const int idx = (threadIdx.x + j * 256) * 2;  // double index within the slab
            const int kk = idx / kHT, mm = idx % kHT;
            const double* src = panel + int64_t(k0 + kk) * Mpad;
            cp_async16(As(st, kk) + mm, src + tile.x * kHT + mm);
            cp_async16(Bs(st, kk) + mm, src + tile.y * kHT + mm);
        }
        cp_async_commit();
    };

    const int nslab = pc_sub / kHK;
    issue(0, 0);
    for (int sl = 0; sl < nslab; ++sl) {
        const int st = sl & 1;
        if (sl + 1 < nslab) {
            issue(st ^ 1, (sl + 1) * kHK);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kHK; ++kk) {
            double a[8], b[8];
            const double* ar = As(st, kk);
            const double* br = Bs(st, kk);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const double2 av = *reinterpret_cast<const double2*>(ar + ty * 2 + 32 * g);
                const double2 bv = *reinterpret_cast<const double2*>(br + tx * 2 + 32 * g);
                a[2 * g] = av.x; a[2 * g + 1] = av.y;
                b[2 * g] = bv.x; b[2 * g + 1] = bv.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();  // slab consumed before it is overwritten two iterations later
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = tile.x * kHT + ty * 2 + 32 * (i >> 1) + (i & 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int col = tile.y * kHT + tx * 2 + 32 * g;
            double2* dst = reinterpret_cast<double2*>(&C[int64_t(row) * Mpad + col]);
            double2 v = *dst;
            v.x += acc[i][2 * g];
            v.y += acc[i][2 * g + 1];
            *dst = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core version of the tile product (mma.sync.m8n8k4.f64 = DMMA.8x8x4): one DMMA does the
// work of eight warp-wide DFMAs with a single issue slot, which is what the DFMA micro-kernel above
// runs out of (FP64 instructions occupy the issue port for two cycles: ncu showed 68.8 % pipe
// utilisation with the math-pipe throttle as top stall, i.e. contraction-bound -- SURVEY.md 8d's
// condition for trying the tensor path; tools/dmma_probe.cu is the measurement behind the choice).
//   block tile 128 x 128, 16 warps as 2 (rows) x 8 (columns), warp tile 64 x 16 = 8 x 2 DMMA tiles; measured on
//   config 4 (M = 1,500, 8.73 M points): 2 x 8 warps 731.7 ms, 2 x 4 744.1 ms, 4 x 4 752.7 ms per Hessian --
//   the layout hardly matters, the tensor pipe is 85 % active either way (profiles/r2_hessian_ncu.txt);
//   per 4-point k-step a warp loads 8 + 2 operand doubles per lane from shared memory for 16 DMMAs;
//   the slab stride is padded to 132 doubles (= 4 mod 16): the fragment pattern
//   (k = lane % 4, column = lane / 4) is then bank-conflict-free.
// Fragment layout (PTX ISA, mma.m8n8k4 .f64): A row-major a0 = A[lane/4][lane%4], B column-major
// b0 = B[lane%4][lane/4], C c0,c1 = C[lane/4][2 (lane%4) + 0,1].  Here A[m][k] = Gu[p0+k][row0+m] and
// B[k][n] = Gu[p0+k][col0+n]: both operands are read from the same point-major slab.
// ---------------------------------------------------------------------------------------------
#ifndef HP_HESS_DK
#define HP_HESS_DK 32
#endif
#ifndef HP_HESS_STAGES
#define HP_HESS_STAGES 3
#endif
constexpr int kDK = HP_HESS_DK;    // points per shared-memory slab (3 stages x 2 x 32 x 132 doubles = 203 KB)
constexpr int kDStride = kHT + 4;  // padded row length of a slab row (doubles)
constexpr int kDStages = HP_HESS_STAGES;
#ifndef HP_HESS_WM
#define HP_HESS_WM 2
#endif
#ifndef HP_HESS_WN
#define HP_HESS_WN 8
#endif
constexpr int kDWM = HP_HESS_WM, kDWN = HP_HESS_WN;  // warps along rows / columns of the 128 x 128 block tile

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // not volatile: a pure function of its operands, so the compiler may hoist the operand loads of the
    // next k-step above the DMMAs of this one
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32)
syrk_panel_dmma_kernel(const double* __restrict__ Gu, int Mpad, int pc_sub, const int2* __restrict__ tiles,
                       double* __restrict__ Cpart, const unsigned long long* __restrict__ flags, int nq) {
    const unsigned live = hessian_live_quadrants(flags, gridDim.y, nq, blockIdx.y, tiles[blockIdx.x]);
    if (!live) return;
    constexpr int kThreads = WM * WN * 32;
    constexpr int TM = kHT / WM / 8, TN = kHT / WN / 8;  // 8 x 8 DMMA tiles per warp along rows / columns
    extern __shared__ __align__(16) double smem_syrk[];  // [stage][A|B][kDK][kDStride]
    auto As = [&](int st, int kk) { return smem_syrk + ((st * 2 + 0) * kDK + kk) * kDStride; };
    auto Bs = [&](int st, int kk) { return smem_syrk + ((st * 2 + 1) * kDK + kk) * kDStride; };
    const int2 tile = tiles[blockIdx.x];
    const int s = blockIdx.y;
    const double* panel = Gu + int64_t(s) * pc_sub * Mpad;
    double* C = Cpart + int64_t(s) * Mpad * Mpad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wrow = (warp / WN) * (8 * TM), wcol = (warp % WN) * (8 * TN);
    static_assert(kHQ % (8 * TM) == 0 && kHQ % (8 * TN) == 0, "a warp's sub-tile must lie inside one quadrant");
    const bool mine = (live >> (2 * (wrow / kHQ) + wcol / kHQ)) & 1u;  // warp-uniform
    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int st, int k0) {
#pragma unroll
        for (int j = 0; j < (kDK * kHT) / (2 * kThreads); ++j) {
            const int idx = (threadIdx.x + j * kThreads) * 2;  // double index within the (unpadded) slab
            const int kk = idx / kHT, mm = idx % kHT;
            const double* src = panel + int64_t(k0 + kk) * Mpad;
            cp_async16(As(st, kk) + mm, src + tile.x * kHT + mm);
            cp_async16(Bs(st, kk) + mm, src + tile.y * kHT + mm);
        }
        cp_async_commit();
    };

    const int nslab = pc_sub / kDK;
    for (int pre = 0; pre < kDStages - 1; ++pre) {
        if (pre < nslab) issue(pre, pre * kDK);
        else cp_async_commit();
    }
    for (int sl = 0; sl < nslab; ++sl) {
        const int st = sl % kDStages;
        cp_async_wait<kDStages - 2>();
        __syncthreads();  // slab `sl` has landed for every thread; slab sl-1 is consumed by everyone
        if (sl + kDStages - 1 < nslab) issue((sl + kDStages - 1) % kDStages, (sl + kDStages - 1) * kDK);
        else cp_async_commit();
        if (!mine) continue;  // this warp's quadrant is screened out: it only feeds the operand pipeline
        // operand fragments of k-step k4 + 4 are loaded while the DMMAs of k-step k4 issue
        double a[2][TM], b[2][TN];
        {
            const double* ar = As(st, tig) + wrow + gid;
            const double* br = Bs(st, tig) + wcol + gid;
#pragma unroll
            for (int i = 0; i < TM; ++i) a[0][i] = ar[8 * i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[0][j] = br[8 * j];
        }
#pragma unroll
        for (int k4 = 0; k4 < kDK; k4 += 4) {
            const int cur = (k4 >> 2) & 1, nxt = cur ^ 1;
            if (k4 + 4 < kDK) {
                const double* ar = As(st, k4 + 4 + tig) + wrow + gid;
                const double* br = Bs(st, k4 + 4 + tig) + wcol + gid;
#pragma unroll
                for (int i = 0; i < TM; ++i) a[nxt][i] = ar[8 * i];
#pragma unroll
                for (int j = 0; j < TN; ++j) b[nxt][j] = br[8 * j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
        }
    }
    cp_async_wait<0>();
    if (!mine) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = tile.x * kHT + wrow + 8 * i + gid;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = tile.y * kHT + wcol + 8 * j + 2 * tig;
            double2* dst = reinterpret_cast<double2*>(&C[int64_t(row) * Mpad + col]);
            double2 v = *dst;
            v.x += acc[i][j][0];
            v.y += acc[i][j][1];
            *dst = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same tile product with the operand slabs moved by the bulk-copy engine (cp.async.bulk = UBLKCP, one
// 1 KB row of a slab per copy, padded destination rows) and an mbarrier ring instead of cp.async groups +
// __syncthreads: a 17th warp is the producer (waits for a stage to be released by all 16 consumer warps,
// arms the stage's `full` barrier with the slab's byte count, issues the 2 x kDK row copies), the consumer
// warps wait for `full`, run their DMMAs and release the stage -- no block-wide barrier in the loop, every
// warp runs at its own pace.  Default; HP_B200_HESSIAN_PIPE=cpasync selects the kernel above.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 16-byte aligned global -> shared bulk copy; completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32 + 32)
syrk_panel_bulk_kernel(const double* __restrict__ Gu, int Mpad, int pc_sub, const int2* __restrict__ tiles,
                       double* __restrict__ Cpart, const unsigned long long* __restrict__ flags, int nq) {
    const unsigned live = hessian_live_quadrants(flags, gridDim.y, nq, blockIdx.y, tiles[blockIdx.x]);
    if (!live) return;
    constexpr int kConsumers = WM * WN;
    constexpr int TM = kHT / WM / 8, TN = kHT / WN / 8;
    constexpr unsigned kRowBytes = kHT * sizeof(double);
    extern __shared__ __align__(16) double smem_syrk[];  // [stage][A|B][kDK][kDStride] | full[stages] | empty[stages]
    auto As = [&](int st, int kk) { return smem_syrk + ((st * 2 + 0) * kDK + kk) * kDStride; };
    auto Bs = [&](int st, int kk) { return smem_syrk + ((st * 2 + 1) * kDK + kk) * kDStride; };
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_syrk + kDStages * 2 * kDK * kDStride);
    unsigned long long* empty = full + kDStages;
    const int2 tile = tiles[blockIdx.x];
    const int s = blockIdx.y;
    const double* panel = Gu + int64_t(s) * pc_sub * Mpad;
    double* C = Cpart + int64_t(s) * Mpad * Mpad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < kDStages; ++st) {
            mbar_init(&full[st], 1);            // the producer's arrive.expect_tx; the bytes complete the phase
            mbar_init(&empty[st], kConsumers);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int nslab = pc_sub / kDK;

    if (warp == kConsumers) {  // ---- producer warp
        for (int sl = 0; sl < nslab; ++sl) {
            const int st = sl % kDStages, round = sl / kDStages;
            if (round > 0) mbar_wait(&empty[st], (round - 1) & 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[st], 2u * kDK * kRowBytes);
            __syncwarp();
            for (int kk = lane; kk < kDK; kk += 32) {
                const double* src = panel + int64_t(sl * kDK + kk) * Mpad;
                bulk_g2s(As(st, kk), src + tile.x * kHT, kRowBytes, &full[st]);
                bulk_g2s(Bs(st, kk), src + tile.y * kHT, kRowBytes, &full[st]);
            }
        }
        return;
    }

    // ---- consumer warps
    const int gid = lane >> 2, tig = lane & 3;
    const int wrow = (warp / WN) * (8 * TM), wcol = (warp % WN) * (8 * TN);
    static_assert(kHQ % (8 * TM) == 0 && kHQ % (8 * TN) == 0, "a warp's sub-tile must lie inside one quadrant");
    const bool mine = (live >> (2 * (wrow / kHQ) + wcol / kHQ)) & 1u;  // warp-uniform
    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int sl = 0; sl < nslab; ++sl) {
        const int st = sl % kDStages, round = sl / kDStages;
        mbar_wait(&full[st], round & 1);
        if (mine) {
            double a[2][TM], b[2][TN];
            {
                const double* ar = As(st, tig) + wrow + gid;
                const double* br = Bs(st, tig) + wcol + gid;
#pragma unroll
                for (int i = 0; i < TM; ++i) a[0][i] = ar[8 * i];
#pragma unroll
                for (int j = 0; j < TN; ++j) b[0][j] = br[8 * j];
            }
#pragma unroll
            for (int k4 = 0; k4 < kDK; k4 += 4) {
                const int cur = (k4 >> 2) & 1, nxt = cur ^ 1;
                if (k4 + 4 < kDK) {
                    const double* ar = As(st, k4 + 4 + tig) + wrow + gid;
                    const double* br = Bs(st, k4 + 4 + tig) + wcol + gid;
#pragma unroll
                    for (int i = 0; i < TM; ++i) a[nxt][i] = ar[8 * i];
#pragma unroll
                    for (int j = 0; j < TN; ++j) b[nxt][j] = br[8 * j];
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);  // this warp has read everything it needs from the stage
    }
    if (!mine) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = tile.x * kHT + wrow + 8 * i + gid;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = tile.y * kHT + wcol + 8 * j + 2 * tig;
            double2* dst = reinterpret_cast<double2*>(&C[int64_t(row) * Mpad + col]);
            double2 v = *dst;
            v.x += acc[i][j][0];
            v.y += acc[i][j][1];
            *dst = v;
        }
    }
}

__global__ void __launch_bounds__(256)
hessian_finish_kernel(int M, int Mpad, int nsplit, const double* __restrict__ Cpart,
                      double* __restrict__ H) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= int64_t(M) * M) return;
    const int m = int(idx / M), n = int(idx % M);
    const int r = m < n ? m : n, c = m < n ? n : m;  // upper triangle holds the data
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += Cpart[int64_t(k) * Mpad * Mpad + int64_t(r) * Mpad + c];
    H[idx] = s;
}

// tiles[k] = (i, j), i <= j, row-major over the upper triangle of the nt x nt tile grid
__global__ void tile_list_kernel(int nt, int2* __restrict__ tiles) {
    for (int k = threadIdx.x; k < nt * (nt + 1) / 2; k += blockDim.x) {
        int i = 0, rem = k;
        while (rem >= nt - i) {
            rem -= nt - i;
            ++i;
        }
        tiles[k] = make_int2(i, i + rem);
    }
}

// Layout of one chunk of the panel and of the scratch buffer.  A chunk is `nsplit` sub-panels of kHSub points
// each; every (tile, sub-panel) pair is one thread block of the tile product and owns its own 128 x 128 slice
// of the partial matrices (no atomics: H is a fixed-order sum of the nsplit partial matrices).  Many short
// sub-panels on purpose: the screening decides per (sub-panel, column block), so short sub-panels (one atom's
// radial shells: 1,280 points are a twentieth of an atomic grid) skip more, and a launch of ntile x nsplit
// blocks of which a fifth is live still fills several waves of SMs.  The panel does not need to fit the L2:
// the blocks of one sub-panel run together (blockIdx.x = tile is the fast index) and share its 3-30 MB.
// Every block ends with a read-modify-write of its slice of the partial matrices, so the chunk is as large
// as the panel budget allows (measured in round 2: 878 ms per Hessian with 64 MB chunks against 738 ms).
constexpr int kHSub = HP_HESS_SUB;                      // points per sub-panel
constexpr int64_t kHPanelBytes = int64_t(HP_HESS_PANEL_MB) << 20;  // per panel buffer (two of them)
static_assert(kHSub % kDK == 0 && kHSub % kHK == 0, "whole slabs of either tile-product kernel");

struct HessianLayout {
    int Mpad, nt, nq, ntile, nsplit, pc_sub, pc;
    size_t panel_doubles, parts_doubles, tile_bytes, nflag;
    size_t bytes() const {
        return 2 * panel_doubles * sizeof(double) + parts_doubles * sizeof(double) + tile_bytes +
               ((2 * nflag * sizeof(unsigned long long) + 255) / 256) * 256;
    }
};

// npts < 0: the largest layout for M (what hp_hessian_scratch_bytes reserves); otherwise no more sub-panels
// than the points need (the offsets inside the scratch buffer then depend on npts, consistently everywhere).
static HessianLayout hessian_layout(int M, int64_t npts) {
    HessianLayout L;
    L.Mpad = ((M + kHT - 1) / kHT) * kHT;
    L.nt = L.Mpad / kHT;
    L.nq = L.Mpad / kHQ;
    L.ntile = L.nt * (L.nt + 1) / 2;
    L.pc_sub = kHSub;
    int64_t ns = kHPanelBytes / (int64_t(kHSub) * L.Mpad * 8);
    ns = ns < 2 ? 2 : (ns > kHMaxSplit ? kHMaxSplit : ns);
    if (npts >= 0) {
        const int64_t need = (npts + kHSub - 1) / kHSub;
        if (need < ns) ns = need < 1 ? 1 : need;
    }
    L.nsplit = int(ns);
    L.pc = L.nsplit * L.pc_sub;
    L.panel_doubles = size_t(L.pc) * L.Mpad;
    L.parts_doubles = size_t(L.nsplit) * L.Mpad * L.Mpad;
    L.tile_bytes = ((size_t(L.ntile) * sizeof(int2) + 255) / 256) * 256;
    L.nflag = size_t(L.nsplit) * L.nq + 2;
    return L;
}

}  // namespace hp

using namespace hp;

extern "C" size_t hp_hessian_scratch_bytes(int32_t M) { return hessian_layout(M, -1).bytes(); }

namespace {

// side stream + events of the panel pipeline, one set per device, created on first use
struct HessianPipe {
    cudaStream_t side = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}, start = nullptr;
};

int hessian_pipe(HessianPipe** out) {
    static HessianPipe pipes[64];
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    HP_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    HessianPipe& p = pipes[dev];
    if (!p.side) {
        rc = check_cuda(cudaStreamCreateWithFlags(&p.side, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < 2 && rc == HP_OK; ++i) {
            rc = check_cuda(cudaEventCreateWithFlags(&p.ready[i], cudaEventDisableTiming), "cudaEventCreate");
            if (rc == HP_OK) rc = check_cuda(cudaEventCreateWithFlags(&p.consumed[i], cudaEventDisableTiming), "cudaEventCreate");
        }
        if (rc == HP_OK) rc = check_cuda(cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming), "cudaEventCreate");
        if (rc) return rc;
    }
    *out = &p;
    return HP_OK;
}

// H = sum over chunks of panel^T panel.  `produce(p0, pc, Mpad, panel, stream)` launches the kernel that fills
// one chunk of the panel.  Panels are double-buffered: chunk c + 1 is generated on a side stream while the
// tile product of chunk c runs (the generator is 6 % of the work and the tile product leaves a third of
// every SM's registers and most issue slots free).
template <class Produce>
int hessian_run(int64_t npts, int32_t M, void* scratch, size_t scratch_bytes, double* H, cudaStream_t st, Produce produce) {
    HP_REQUIRE(scratch_bytes >= hp_hessian_scratch_bytes(M), "scratch too small");
    const HessianLayout L = hessian_layout(M, npts);
    const int Mpad = L.Mpad, nt = L.nt, nq = L.nq, ntile = L.ntile, nsplit = L.nsplit, pc = L.pc, pc_sub = L.pc_sub;
    double* panel[2] = {static_cast<double*>(scratch), static_cast<double*>(scratch) + L.panel_doubles};
    double* parts = panel[1] + L.panel_doubles;
    int2* tiles = reinterpret_cast<int2*>(parts + L.parts_doubles);
    const size_t nflag = L.nflag;
    unsigned long long* flag_base = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(tiles) + L.tile_bytes);
    const char* screen_env = getenv("HP_B200_HESSIAN_SCREEN");  // read per call: tests compare both settings
    const bool screen = !(screen_env && screen_env[0] == '0');
    unsigned long long* flags[2] = {screen ? flag_base : nullptr, screen ? flag_base + nflag : nullptr};
    // tensor-core (DMMA) tile product by default; HP_B200_HESSIAN_DFMA=1 selects the vector-FMA kernel
    static const bool use_dfma = [] { const char* e = getenv("HP_B200_HESSIAN_DFMA"); return e && e[0] == '1'; }();
    // operand pipeline of the DMMA kernel: bulk copies + mbarriers (default) or cp.async groups + __syncthreads
    const char* pipe_env = getenv("HP_B200_HESSIAN_PIPE");  // read per call: a test compares the two bit for bit
    const bool use_cpasync = pipe_env && pipe_env[0] == 'c';
    constexpr size_t kBulkSmem = sizeof(double) * kDStages * 2 * kDK * kDStride + 2 * kDStages * sizeof(unsigned long long);
    const size_t syrk_smem = use_dfma ? sizeof(double) * 2 * 2 * kHK * kHT
                                      : (use_cpasync ? sizeof(double) * kDStages * 2 * kDK * kDStride : kBulkSmem);
    {
        static bool configured[64] = {};  // attributes are per function and device
        if (first_use_on_device(configured)) {
            int rc0 = check_cuda(cudaFuncSetAttribute(syrk_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      int(sizeof(double) * 2 * 2 * kHK * kHT)), "cudaFuncSetAttribute");
            if (rc0) return rc0;
            rc0 = check_cuda(cudaFuncSetAttribute(syrk_panel_dmma_kernel<kDWM, kDWN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  int(sizeof(double) * kDStages * 2 * kDK * kDStride)), "cudaFuncSetAttribute");
            if (rc0) return rc0;
            rc0 = check_cuda(cudaFuncSetAttribute(syrk_panel_bulk_kernel<kDWM, kDWN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  int(kBulkSmem)), "cudaFuncSetAttribute");
            if (rc0) return rc0;
        }
    }
    HessianPipe* pipe = nullptr;
    int rc = hessian_pipe(&pipe);
    if (rc) return rc;
    // tile list (upper triangle), written on the device: no host allocation, copy or synchronisation
    tile_list_kernel<<<1, 256, 0, st>>>(nt, tiles);
    HP_LAUNCH_CHECK("tile_list_kernel");
    rc = check_cuda(cudaMemsetAsync(parts, 0, sizeof(double) * nsplit * size_t(Mpad) * Mpad, st), "memset");
    if (rc == HP_OK) rc = check_cuda(cudaMemsetAsync(flag_base, 0, sizeof(unsigned long long) * 2 * nflag, st), "memset");
    if (rc) return rc;
    // the side stream starts after everything already queued on `st` (promolecule, weights)
    rc = check_cuda(cudaEventRecord(pipe->start, st), "cudaEventRecord");
    if (rc == HP_OK) rc = check_cuda(cudaStreamWaitEvent(pipe->side, pipe->start, 0), "cudaStreamWaitEvent");
    int64_t chunk = 0;
    for (int64_t p0 = 0; p0 < npts && rc == HP_OK; p0 += pc, ++chunk) {
        const int b = int(chunk & 1);
        if (chunk >= 2) rc = check_cuda(cudaStreamWaitEvent(pipe->side, pipe->consumed[b], 0), "cudaStreamWaitEvent");
        // executed-tile counter (last slot) accumulates over the chunks of one buffer: zero the maxima only
        if (rc == HP_OK && screen)
            rc = check_cuda(cudaMemsetAsync(flags[b], 0, sizeof(unsigned long long) * (nflag - 1), pipe->side), "memset");
        if (rc == HP_OK) rc = produce(p0, pc, Mpad, panel[b], flags[b], pc_sub, pipe->side);
        if (rc == HP_OK && screen) {
            panel_chunk_max_kernel<<<1, 256, 0, pipe->side>>>(nsplit * nq, flags[b]);
            HP_LAUNCH_CHECK("panel_chunk_max_kernel");
        }
        if (rc == HP_OK) rc = check_cuda(cudaEventRecord(pipe->ready[b], pipe->side), "cudaEventRecord");
        if (rc == HP_OK) rc = check_cuda(cudaStreamWaitEvent(st, pipe->ready[b], 0), "cudaStreamWaitEvent");
        if (rc) break;
        if (use_dfma) syrk_panel_kernel<<<dim3(ntile, nsplit), 256, syrk_smem, st>>>(panel[b], Mpad, pc_sub, tiles, parts, flags[b], nq);
        else if (use_cpasync) syrk_panel_dmma_kernel<kDWM, kDWN><<<dim3(ntile, nsplit), kDWM * kDWN * 32, syrk_smem, st>>>(panel[b], Mpad, pc_sub, tiles, parts, flags[b], nq);
        else syrk_panel_bulk_kernel<kDWM, kDWN><<<dim3(ntile, nsplit), kDWM * kDWN * 32 + 32, syrk_smem, st>>>(panel[b], Mpad, pc_sub, tiles, parts, flags[b], nq);
        HP_LAUNCH_CHECK("syrk_panel_kernel");
        rc = check_cuda(cudaEventRecord(pipe->consumed[b], st), "cudaEventRecord");
    }
    if (rc) return rc;
    const int64_t total = int64_t(M) * M;
    hessian_finish_kernel<<<int((total + 255) / 256), 256, 0, st>>>(M, Mpad, nsplit, parts, H);
    HP_LAUNCH_CHECK("hessian_finish_kernel");
    return HP_OK;
}

}  // namespace

extern "C" int hp_hessian(int functor, int64_t npts, const double* px, const double* py,
                          const double* pz, const double* atom_xyz, const int32_t* shell_atom,
                          const double* shell_norm, const double* shell_alpha,
                          const double* shell_order, const double* rho, const double* molw,
                          const double* promol, double density_cutoff, int32_t M, void* scratch,
                          size_t scratch_bytes, double* H, void* stream) {
    HP_REQUIRE(npts > 0 && M > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && shell_atom && shell_norm && shell_alpha && rho && molw &&
                   promol && scratch && H, "null input");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    HP_REQUIRE(functor == HP_FUNCTOR_SLATER || functor == HP_FUNCTOR_GAUSS || functor == HP_FUNCTOR_GENERAL,
               "unsupported functor");
    auto produce = [&](int64_t p0, int pc, int Mpad, double* panel, unsigned long long* flags, int pc_sub,
                       cudaStream_t s) -> int {
#define HP_BASIS(F)                                                                                         \
    basis_chunk_kernel<F><<<pc / kHRows, 256, 0, s>>>(p0, pc, M, Mpad, px, py, pz, rho, molw, promol, density_cutoff, \
                                             shell_atom, atom_xyz, shell_norm, shell_alpha, shell_order, npts, panel, \
                                             flags, pc_sub)
        if (functor == HP_FUNCTOR_SLATER) HP_BASIS(HP_FUNCTOR_SLATER);
        else if (functor == HP_FUNCTOR_GAUSS) HP_BASIS(HP_FUNCTOR_GAUSS);
        else HP_BASIS(HP_FUNCTOR_GENERAL);
#undef HP_BASIS
        HP_LAUNCH_CHECK("basis_chunk_kernel");
        return HP_OK;
    };
    return hessian_run(npts, M, scratch, scratch_bytes, H, as_stream(stream), produce);
}

extern "C" int hp_hessian_table(int64_t npts, const double* px, const double* py, const double* pz,
                                const double* atom_xyz, const int32_t* shell_atom, const int32_t* knot_offsets,
                                const double* knots, const int32_t* lut_meta, const uint16_t* lut,
                                const int64_t* shell_coef_offsets, const double* shell_coef, const double* rho,
                                const double* molw, const double* promol, double density_cutoff, int32_t M,
                                void* scratch, size_t scratch_bytes, double* H, void* stream) {
    HP_REQUIRE(npts > 0 && M > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && shell_atom && knot_offsets && knots && lut_meta && lut &&
                   shell_coef_offsets && shell_coef && rho && molw && promol && scratch && H, "null input");
    TableArgs tab{knot_offsets, knots, lut_meta, lut, reinterpret_cast<const long long*>(shell_coef_offsets), shell_coef};
    auto produce = [&](int64_t p0, int pc, int Mpad, double* panel, unsigned long long* flags, int pc_sub,
                       cudaStream_t s) -> int {
        table_basis_chunk_kernel<<<pc / kHRows, 256, 0, s>>>(p0, pc, M, Mpad, px, py, pz, rho, molw, promol, density_cutoff,
                                                    shell_atom, atom_xyz, tab, npts, panel, flags, pc_sub);
        HP_LAUNCH_CHECK("table_basis_chunk_kernel");
        return HP_OK;
    };
    return hessian_run(npts, M, scratch, scratch_bytes, H, as_stream(stream), produce);
}

// 64 x 64 quadrant products executed by the last hp_hessian / hp_hessian_table call that used `scratch` (sum
// over chunks and sub-panels; one quadrant = 64 x 64 x *points_per_tile_out multiply-adds) and the number the
// unscreened product needs (four per tile, three for the tiles on the diagonal).  Synchronises `stream`.
// executed = 0 when screening is disabled.
extern "C" int hp_hessian_tiles_executed(int32_t M, int64_t npts, const void* scratch, int64_t* executed_out,
                                         int64_t* total_out, int32_t* points_per_tile_out, void* stream) {
    HP_REQUIRE(M > 0 && npts > 0 && scratch && executed_out && total_out && points_per_tile_out, "bad arguments");
    const HessianLayout L = hessian_layout(M, npts);
    const char* base = static_cast<const char*>(scratch) + (2 * L.panel_doubles + L.parts_doubles) * sizeof(double) + L.tile_bytes;
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(base);
    unsigned long long host[2] = {0, 0};
    cudaStream_t st = as_stream(stream);
    int rc = check_cuda(cudaMemcpyAsync(&host[0], f + L.nflag - 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "copy");
    if (rc == HP_OK) rc = check_cuda(cudaMemcpyAsync(&host[1], f + 2 * L.nflag - 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "copy");
    if (rc == HP_OK) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    if (rc) return rc;
    const int64_t nchunk = (npts + L.pc - 1) / L.pc;
    *executed_out = int64_t(host[0] + host[1]);
    *total_out = nchunk * L.nsplit * (int64_t(4) * L.ntile - L.nt);
    *points_per_tile_out = L.pc_sub;
    return HP_OK;
}
