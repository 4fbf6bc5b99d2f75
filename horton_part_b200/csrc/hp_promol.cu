// Fused promolecule / owner-weight / entropy pass (rows a1-a7, a10 of SURVEY.md section 8a).
//
// Work decomposition: one thread owns kPts grid points (strided by blockDim so loads coalesce),
// keeps their running promolecule sums in registers, and walks over ALL atoms in reference order.
// Atoms are staged through shared memory in tiles (coordinates + interleaved (A, alpha) shell
// table); every thread of the block reads the same atom at the same time, so the shared-memory
// reads are broadcasts.  Distances are recomputed in registers: the reference's cached
// natom x Npts `radial_distances` table (core/base.py:630-635) is never materialised.
//
// Bound: the FP64 pipe (there is no FP64 SFU path for exp): ~38 FP64 instructions per
// atom x point evaluation at K = 4/3 shells, against 56 B of HBM traffic per *point*.  The inner
// loop is written so that almost every issued instruction is an FP64 FMA:
//   * sqrt and exp are inlined without library slow-path calls (MUFU.RSQ64H seed + Goldschmidt and
//     Heron steps; Cody-Waite reduction + degree-11 polynomial with constant-bank coefficients);
//   * the owner atom's pro-atom value is recomputed once per point after the loop instead of a
//     compare/select per pair;
//   * exp underflow (arg <= -708) is flushed to zero with two selects instead of a branch.
#include "hp_common.cuh"
#include "hp_math.cuh"
#include "hp_promol_common.cuh"

// Tuning notes (B200, config 5, ms per iteration): library sqrt()/exp() with 2 points/thread 491;
// inlined no-call sqrt/exp, 4 points/thread 316; 6-op sqrt 305; first shell peeled + next atom
// prefetched from shared memory one iteration ahead ~293.  Table-driven exp variants (16/32/64
// entries in shared memory, 11-13 FP64 ops instead of 16) were slower (362/352/346): the per-lane
// 8-byte shared loads and index arithmetic cost more issue slots than the DFMAs they save.
#ifndef HP_PTS
#define HP_PTS 4
#endif
#ifndef HP_THREADS
#define HP_THREADS 256
#endif
#ifndef HP_MINBLOCKS
#define HP_MINBLOCKS 2
#endif

namespace hp {

constexpr int kThreads = HP_THREADS;
constexpr int kPts = HP_PTS;            // points per thread

template <int F>
__global__ void __launch_bounds__(kThreads, HP_MINBLOCKS)
promol_weights_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                      const double* __restrict__ pz, int64_t point_base, int natom,
                      const double* __restrict__ atom_xyz, const int64_t* __restrict__ atom_pt_off,
                      const int* __restrict__ atom_sh_off, const double* __restrict__ shell_A,
                      const double* __restrict__ shell_alpha, const double* __restrict__ shell_order,
                      int ntile, const int* __restrict__ tile_off, const double* __restrict__ rho,
                      const double* __restrict__ molw, double density_cutoff, double promol_offset,
                      double* __restrict__ promol_out, double* __restrict__ w_out,
                      double* __restrict__ entropy_partials) {
    __shared__ AtomRec s_atoms[kTileAtoms + 1];  // +1: sentinel for the prefetch
    __shared__ double2 s_AB[kTileShells];
    __shared__ double s_N[(F == HP_FUNCTOR_GENERAL) ? kTileShells : 1];
    __shared__ double s_red[32];

    const int64_t span = int64_t(kThreads) * kPts;
    const int64_t nchunk = (npts + span - 1) / span;
    double entropy_acc = 0.0;

    for (int64_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        double x[kPts], y[kPts], z[kPts], pro[kPts];
        int64_t q[kPts];
#pragma unroll
        for (int j = 0; j < kPts; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kThreads + threadIdx.x;
            q[j] = p < npts ? p : (npts - 1);  // tail lanes recompute the last point, never store
            x[j] = px[q[j]];
            y[j] = py[q[j]];
            z[j] = pz[q[j]];
            pro[j] = 0.0;
        }

        for (int t = 0; t < ntile; ++t) {
            const int a0 = tile_off[t], a1 = tile_off[t + 1];
            const int sh0 = atom_sh_off[a0], sh1 = atom_sh_off[a1];
            __syncthreads();  // previous tile fully consumed
            for (int i = threadIdx.x; i < a1 - a0; i += kThreads) {
                AtomRec rec;
                rec.x = atom_xyz[3 * (a0 + i) + 0];
                rec.y = atom_xyz[3 * (a0 + i) + 1];
                rec.z = atom_xyz[3 * (a0 + i) + 2];
                rec.s0 = atom_sh_off[a0 + i] - sh0;
                rec.ns = atom_sh_off[a0 + i + 1] - atom_sh_off[a0 + i];
                s_atoms[i] = rec;
            }
            if (threadIdx.x == 0) {
                AtomRec sentinel = {0.0, 0.0, 0.0, 0, 0};
                s_atoms[a1 - a0] = sentinel;
            }
            for (int i = threadIdx.x; i < sh1 - sh0; i += kThreads) {
                s_AB[i] = make_double2(shell_A[sh0 + i], shell_alpha[sh0 + i]);
                if (F == HP_FUNCTOR_GENERAL) s_N[i] = shell_order[sh0 + i];
            }
            __syncthreads();

            // software pipelining: the next atom's record and first shell are fetched from shared
            // memory one iteration ahead so their latency never sits on the FP64 critical path
            AtomRec nxt = s_atoms[0];
            double2 nxt_ab = s_AB[nxt.s0];
            for (int i = 0; i < a1 - a0; ++i) {
                const AtomRec rec = nxt;
                const double2 ab0 = nxt_ab;
                nxt = s_atoms[i + 1];          // entry [a1-a0] is a sentinel
                nxt_ab = s_AB[nxt.s0];
                double d2[kPts], f[kPts];
#pragma unroll
                for (int j = 0; j < kPts; ++j) {
                    const double dx = x[j] - rec.x, dy = y[j] - rec.y, dz = z[j] - rec.z;
                    d2[j] = fma(dz, dz, fma(dy, dy, dx * dx));
                }
                eval_proatom<F, kPts>(d2, rec.s0, rec.ns, s_AB, s_N, f, ab0);
                // update_pro, core/stockholder.py:169-170: promoldens += work; += 1e-100
#pragma unroll
                for (int j = 0; j < kPts; ++j) pro[j] = (pro[j] + f[j]) + promol_offset;
            }
        }

        // Epilogue per point: owner atom's pro-atom (recomputed with the same code path, hence
        // bit-identical to the value that entered the sum), weight, entropy term.
#pragma unroll
        for (int j = 0; j < kPts; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kThreads + threadIdx.x;
            if (p >= npts) continue;
            if (promol_out) promol_out[p] = pro[j];
            if (w_out) {
                const int64_t g = point_base + p;
                int lo = 0, hi = natom;  // invariant: atom_pt_off[lo] <= g < atom_pt_off[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (atom_pt_off[mid] <= g) lo = mid; else hi = mid;
                }
                const double dx = x[j] - atom_xyz[3 * lo + 0], dy = y[j] - atom_xyz[3 * lo + 1],
                             dz = z[j] - atom_xyz[3 * lo + 2];
                double d2[1] = {fma(dz, dz, fma(dy, dy, dx * dx))}, own[1];
                const int s0 = atom_sh_off[lo], ns = atom_sh_off[lo + 1] - s0;
                // shell parameters straight from global memory (same values as the tile copy)
                double fo = 0.0;
                {
                    const double r = (F == HP_FUNCTOR_GAUSS) ? d2[0] : sqrt_nocall(d2[0]);
                    for (int k = 0; k < ns; ++k) {
                        const double A = shell_A[s0 + k], al = shell_alpha[s0 + k];
                        if (F == HP_FUNCTOR_GENERAL) {
                            const double n = shell_order[s0 + k];
                            const double rn = (n == 1.0) ? r : ((n == 2.0) ? r * r : pow(r, n));
                            fo = fma(A, exp(-al * rn), fo);
                        } else {
                            fo = fma(A, exp_neg_poly(-al * r), fo);
                        }
                    }
                }
                own[0] = fo;
                // core/stockholder.py:376-377: w /= promol; clip to [0, 1]
                double w = own[0] / pro[j];
                w = fmin(fmax(w, 0.0), 1.0);
                w_out[p] = w;
            }
            if (entropy_partials) {
                // core/stockholder.py:145-151: masked rho*ln(rho/rho0), integrated with mol weights
                const double r = rho[p];
                const bool sick = (pro[j] < density_cutoff) || (r < density_cutoff);
                if (!sick) entropy_acc += molw[p] * r * log(r / pro[j]);
            }
        }
    }

    if (entropy_partials) {
        const double total = block_sum(entropy_acc, s_red);
        if (threadIdx.x == 0) entropy_partials[blockIdx.x] = total;
        // blocks beyond the launched grid: zero-filled by block 0
        if (blockIdx.x == 0)
            for (int i = gridDim.x + threadIdx.x; i < kMaxPartials; i += kThreads) entropy_partials[i] = 0.0;
    }
}

static int promol_grid(int64_t npts, const void* kernel) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    const int64_t span = int64_t(kThreads) * kPts;
    int64_t want = (npts + span - 1) / span;
    int64_t cap = int64_t(sm_count()) * per_sm;
    if (cap > kMaxPartials) cap = kMaxPartials;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return int(want);
}

}  // namespace hp

extern "C" int32_t hp_num_partials(void) { return hp::kMaxPartials; }

extern "C" void hp_tile_limits(int32_t* max_atoms_host, int32_t* max_shells_host) {
    *max_atoms_host = hp::kTileAtoms;
    *max_shells_host = hp::kTileShells;
}

extern "C" int hp_promol_weights(int functor, int64_t npts, const double* px, const double* py,
                                 const double* pz, int64_t point_base, int32_t natom,
                                 const double* atom_xyz, const int64_t* atom_point_offsets,
                                 const int32_t* atom_shell_offsets, const double* shell_A,
                                 const double* shell_alpha, const double* shell_order,
                                 int32_t ntile, const int32_t* tile_atom_offsets,
                                 const double* rho, const double* molw, double density_cutoff,
                                 double promol_offset, double* promol, double* at_weights,
                                 double* entropy_partials, void* stream) {
    using namespace hp;
    HP_REQUIRE(npts >= 0 && natom > 0 && ntile > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_point_offsets && atom_shell_offsets, "null input");
    HP_REQUIRE(shell_A && shell_alpha && tile_atom_offsets, "null shell table");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    HP_REQUIRE(!entropy_partials || (rho && molw), "entropy needs rho and molw");
    if (npts == 0) {
        if (entropy_partials)
            return check_cuda(cudaMemsetAsync(entropy_partials, 0, sizeof(double) * kMaxPartials,
                                              as_stream(stream)), "memset partials");
        return HP_OK;
    }
#define HP_LAUNCH_PROMOL(F)                                                                          \
    {                                                                                                \
        const int grid = promol_grid(npts, (const void*)promol_weights_kernel<F>);                   \
        promol_weights_kernel<F><<<grid, kThreads, 0, as_stream(stream)>>>(                          \
            npts, px, py, pz, point_base, natom, atom_xyz, atom_point_offsets, atom_shell_offsets,   \
            shell_A, shell_alpha, shell_order, ntile, tile_atom_offsets, rho, molw, density_cutoff,  \
            promol_offset, promol, at_weights, entropy_partials);                                                \
    }
    switch (functor) {
        case HP_FUNCTOR_SLATER: HP_LAUNCH_PROMOL(HP_FUNCTOR_SLATER); break;
        case HP_FUNCTOR_GAUSS: HP_LAUNCH_PROMOL(HP_FUNCTOR_GAUSS); break;
        case HP_FUNCTOR_GENERAL: HP_LAUNCH_PROMOL(HP_FUNCTOR_GENERAL); break;
        default: set_error("hp_promol_weights: unsupported functor %d", functor); return HP_ERR_ARG;
    }
#undef HP_LAUNCH_PROMOL
    HP_LAUNCH_CHECK("promol_weights_kernel");
    return HP_OK;
}
