// Fused promolecule / owner-weight / entropy pass (rows a1-a7, a10 of SURVEY.md section 8a).
//
// Work decomposition: one thread owns kPts grid points (strided by blockDim so loads coalesce),
// keeps their running promolecule sums in registers, and walks over ALL atoms in reference order.
// Atoms are staged through shared memory in tiles (coordinates + shell table), every thread of the
// block reads the same atom at the same time, so the shared-memory reads are broadcasts.  Distances
// are recomputed in registers: the reference's cached natom x Npts `radial_distances` table
// (core/base.py:630-635) is never materialised.
//
// Bound: FP64 pipe (no FP64 SFU path for exp): ~16 + 36 K flop per atom x point evaluation for
// Slater shells against 48 B of HBM traffic per *point*.
#include "hp_common.cuh"

namespace hp {

constexpr int kThreads = 256;
constexpr int kPts = 2;                 // points per thread
constexpr int kTileAtoms = 128;         // atoms per shared-memory tile
constexpr int kTileShells = 1024;       // shells per shared-memory tile
constexpr int kMaxPartials = 4096;      // size of the entropy partial-sum buffer

struct __align__(16) AtomRec {
    double x, y, z;
    int s0, ns;  // first shell (tile-relative) and shell count
};

template <int F>
__device__ __forceinline__ double eval_proatom(double d2, int s0, int ns, const double* __restrict__ sA,
                                               const double* __restrict__ sAl,
                                               const double* __restrict__ sN) {
    double y = 0.0;
    if (F == HP_FUNCTOR_GAUSS) {
        for (int k = 0; k < ns; ++k) y += sA[s0 + k] * exp(-sAl[s0 + k] * d2);
    } else if (F == HP_FUNCTOR_SLATER) {
        const double r = sqrt(d2);
        for (int k = 0; k < ns; ++k) y += sA[s0 + k] * exp(-sAl[s0 + k] * r);
    } else {
        const double r = sqrt(d2);
        for (int k = 0; k < ns; ++k) {
            const double n = sN[s0 + k];
            const double rn = (n == 1.0) ? r : ((n == 2.0) ? r * r : pow(r, n));
            y += sA[s0 + k] * exp(-sAl[s0 + k] * rn);
        }
    }
    return y;
}

template <int F>
__global__ void __launch_bounds__(kThreads)
promol_weights_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                      const double* __restrict__ pz, int64_t point_base, int natom,
                      const double* __restrict__ atom_xyz, const int64_t* __restrict__ atom_pt_off,
                      const int* __restrict__ atom_sh_off, const double* __restrict__ shell_A,
                      const double* __restrict__ shell_alpha, const double* __restrict__ shell_order,
                      int ntile, const int* __restrict__ tile_off, const double* __restrict__ rho,
                      const double* __restrict__ molw, double density_cutoff,
                      double* __restrict__ promol_out, double* __restrict__ w_out,
                      double* __restrict__ entropy_partials) {
    __shared__ AtomRec s_atoms[kTileAtoms];
    __shared__ double s_A[kTileShells];
    __shared__ double s_Al[kTileShells];
    __shared__ double s_N[(F == HP_FUNCTOR_GENERAL) ? kTileShells : 1];
    __shared__ double s_red[32];

    const int64_t span = int64_t(kThreads) * kPts;
    const int64_t nchunk = (npts + span - 1) / span;
    double entropy_acc = 0.0;

    for (int64_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        double x[kPts], y[kPts], z[kPts], pro[kPts], own[kPts];
        int owner[kPts];
        bool live[kPts];
#pragma unroll
        for (int j = 0; j < kPts; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kThreads + threadIdx.x;
            live[j] = p < npts;
            const int64_t q = live[j] ? p : (npts - 1);
            x[j] = px[q];
            y[j] = py[q];
            z[j] = pz[q];
            pro[j] = 0.0;
            own[j] = 0.0;
            // owner = last atom whose first point is <= global index (empty slices are skipped)
            const int64_t g = point_base + q;
            int lo = 0, hi = natom;  // invariant: atom_pt_off[lo] <= g < atom_pt_off[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (atom_pt_off[mid] <= g) lo = mid; else hi = mid;
            }
            owner[j] = lo;
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
            for (int i = threadIdx.x; i < sh1 - sh0; i += kThreads) {
                s_A[i] = shell_A[sh0 + i];
                s_Al[i] = shell_alpha[sh0 + i];
                if (F == HP_FUNCTOR_GENERAL) s_N[i] = shell_order[sh0 + i];
            }
            __syncthreads();

            for (int i = 0; i < a1 - a0; ++i) {
                const AtomRec rec = s_atoms[i];
                const int a = a0 + i;
#pragma unroll
                for (int j = 0; j < kPts; ++j) {
                    const double dx = x[j] - rec.x, dy = y[j] - rec.y, dz = z[j] - rec.z;
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    const double f = eval_proatom<F>(d2, rec.s0, rec.ns, s_A, s_Al, s_N);
                    // update_pro, core/stockholder.py:169-170: promoldens += work; += 1e-100
                    pro[j] = (pro[j] + f) + 1e-100;
                    if (a == owner[j]) own[j] = f;
                }
            }
        }

#pragma unroll
        for (int j = 0; j < kPts; ++j) {
            if (!live[j]) continue;
            const int64_t p = chunk * span + int64_t(j) * kThreads + threadIdx.x;
            if (promol_out) promol_out[p] = pro[j];
            if (w_out) {
                // core/stockholder.py:376-377: w /= promol; clip to [0, 1]
                double w = own[j] / pro[j];
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
                                 double* promol, double* at_weights, double* entropy_partials,
                                 void* stream) {
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
            promol, at_weights, entropy_partials);                                                   \
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
