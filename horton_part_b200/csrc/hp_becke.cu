// Becke fuzzy-cell weights on the device (grid construction becomes the
// bottleneck at 2,000 atoms: O(natom^2 Npts) in qc-grid's BeckeWeights; call sites
// scripts/generate_density.py:102-111, becke.py:107-116).
//
//   w(p) = P_o(p) / sum_a P_a(p),   P_a = prod_{b != a} s(nu_ab),
//   mu_ab = (|p-R_a| - |p-R_b|) / R_ab,  nu_ab = mu_ab + a_ab (1 - mu_ab^2),
//   s(nu) = (1 - f^order(nu)) / 2,  f(x) = 1.5 x - 0.5 x^3,  a_ab from the Bragg-Slater radii ratio,
//   clipped to +-0.45 (table computed by the caller).
//
// One thread per grid point.  Exactness-preserving pruning: the cell function of an atom a that
// loses against the point's NEAREST atom c by more than 2^-80 (s_ac < 2^-80) cannot change the
// FP64 denominator (P_a <= s_ac, while P_c is a product of a handful of factors >= ~1/2), so P_a is
// only accumulated for the few atoms that survive this O(natom) screen; each surviving P_a is the
// full product over all b.  Cost O((1 + survivors) natom) per point instead of O(natom^2).
#include "hp_common.cuh"
#include "hp_math.cuh"

namespace hp {

constexpr int kBkThreads = 128;

__device__ __forceinline__ double becke_switch(double mu, double a, int order) {
    double nu = fma(a, fma(-mu, mu, 1.0), mu);
    for (int i = 0; i < order; ++i) nu = nu * fma(-0.5 * nu, nu, 1.5);
    return 0.5 * (1.0 - nu);
}

__global__ void __launch_bounds__(kBkThreads)
becke_weights_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                     const double* __restrict__ pz, int64_t point_base, int natom,
                     const double* __restrict__ atom_xyz, const int64_t* __restrict__ atom_pt_off,
                     const double* __restrict__ inv_rab, const double* __restrict__ aab, int order,
                     double* __restrict__ out) {
    extern __shared__ double s_xyz[];  // 3 * natom
    for (int i = threadIdx.x; i < 3 * natom; i += kBkThreads) s_xyz[i] = atom_xyz[i];
    __syncthreads();
    const int64_t p = int64_t(blockIdx.x) * kBkThreads + threadIdx.x;
    if (p >= npts) return;
    const double x = px[p], y = py[p], z = pz[p];
    auto dist = [&](int a) {
        const double dx = x - s_xyz[3 * a], dy = y - s_xyz[3 * a + 1], dz = z - s_xyz[3 * a + 2];
        return sqrt_nocall(fma(dz, dz, fma(dy, dy, dx * dx)));
    };
    // owner and nearest atom
    const int64_t g = point_base + p;
    int lo = 0, hi = natom;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (atom_pt_off[mid] <= g) lo = mid; else hi = mid;
    }
    const int owner = lo;
    int c = 0;
    double nc = dist(0);
    for (int a = 1; a < natom; ++a) {
        const double d = dist(a);
        if (d < nc) {
            nc = d;
            c = a;
        }
    }
    double sum = 0.0, p_owner = 0.0;
    for (int a = 0; a < natom; ++a) {
        const double na = dist(a);
        if (a != c && a != owner) {
            const double s_ac = becke_switch((na - nc) * inv_rab[int64_t(a) * natom + c],
                                             aab[int64_t(a) * natom + c], order);
            if (s_ac < 8.271806125530277e-25) continue;  // 2^-80
        }
        const double* ir = inv_rab + int64_t(a) * natom;
        const double* aa = aab + int64_t(a) * natom;
        double prod = 1.0;
        for (int b = 0; b < natom; ++b) {
            if (b == a) continue;
            prod *= becke_switch((na - dist(b)) * ir[b], aa[b], order);
        }
        sum += prod;
        if (a == owner) p_owner = prod;
    }
    out[p] = p_owner / sum;
}

}  // namespace hp

using namespace hp;

extern "C" int hp_becke_weights(int64_t npts, const double* px, const double* py, const double* pz,
                                int64_t point_base, int32_t natom, const double* atom_xyz,
                                const int64_t* atom_point_offsets, const double* inv_rab,
                                const double* aab, int32_t order, double* out, void* stream) {
    HP_REQUIRE(npts >= 0 && natom > 0 && order >= 1, "bad sizes");
    if (npts == 0) return HP_OK;
    HP_REQUIRE(px && py && pz && atom_xyz && atom_point_offsets && inv_rab && aab && out, "null input");
    const size_t smem = sizeof(double) * 3 * size_t(natom);
    HP_REQUIRE(smem <= 200 * 1024, "too many atoms for the shared-memory coordinate table");
    if (smem > 48 * 1024) {
        int rc = check_cuda(cudaFuncSetAttribute(becke_weights_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                            "cudaFuncSetAttribute");
        if (rc) return rc;
    }
    const int64_t blocks = (npts + kBkThreads - 1) / kBkThreads;
    HP_REQUIRE(blocks < (int64_t(1) << 31), "grid too large for one launch");
    becke_weights_kernel<<<int(blocks), kBkThreads, smem, as_stream(stream)>>>(
        npts, px, py, pz, point_base, natom, atom_xyz, atom_point_offsets, inv_rab, aab, order, out);
    HP_LAUNCH_CHECK("becke_weights_kernel");
    return HP_OK;
}
