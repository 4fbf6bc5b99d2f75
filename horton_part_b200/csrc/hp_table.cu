// gLISA with tabulated basis functions (basis_type="numeric"): the shell integrals
//     I_m = sum_p t(p) S_m(|r_p - R_a(m)|),   t = molw * rho / rho0^power  (masked),
// i.e. function_g and the gradient (glisa.py:850-879, 454-458) with S_m a spline instead of an exponential.
// One thread block walks chunks of points; per atom every thread looks the interval up once for each of its
// points and evaluates the atom's shells from their coefficient blocks (L1 / L2 resident: K x 4.8 KB per
// element).  Per-block partial sums, folded in a fixed order: bit-reproducible.  This path is for parity
// with the reference's feature, not a tuned kernel: tabulated bases are a niche of gLISA.
#include "hp_math.cuh"
#include "hp_table.cuh"

namespace hp {

constexpr int kTbThreads = 256;
constexpr int kTbWarps = kTbThreads / 32;
constexpr int kTbPts = 2;
constexpr int kTbMaxShells = 64;  // shells per atom

__global__ void __launch_bounds__(kTbThreads)
table_moments_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                     const double* __restrict__ pz, int natom, const double* __restrict__ atom_xyz,
                     const int* __restrict__ atom_sh_off, TableArgs tab, const double* __restrict__ rho,
                     const double* __restrict__ molw, const double* __restrict__ promol, double density_cutoff,
                     int power, int nshell, double* __restrict__ partial) {
    __shared__ double s_acc[kTbWarps][kTbMaxShells];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* my_partial = partial + int64_t(blockIdx.x) * nshell;
    for (int i = threadIdx.x; i < nshell; i += kTbThreads) my_partial[i] = 0.0;
    const int64_t span = int64_t(kTbThreads) * kTbPts;
    const int64_t nchunk = (npts + span - 1) / span;
    for (int64_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        double x[kTbPts], y[kTbPts], z[kTbPts], t[kTbPts];
#pragma unroll
        for (int j = 0; j < kTbPts; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kTbThreads + threadIdx.x;
            const bool live = p < npts;
            const int64_t q = live ? p : npts - 1;
            x[j] = px[q]; y[j] = py[q]; z[j] = pz[q];
            const double r0 = promol[q], rh = rho[q];
            const bool sick = (rh < density_cutoff) || (r0 < density_cutoff);
            double v = sick ? 0.0 : molw[q] * rh / r0;
            if (power == 2 && !sick) v /= r0;
            t[j] = live ? v : 0.0;
        }
        for (int a = 0; a < natom; ++a) {
            const double ax = atom_xyz[3 * a], ay = atom_xyz[3 * a + 1], az = atom_xyz[3 * a + 2];
            const int sh0 = atom_sh_off[a], ns = atom_sh_off[a + 1] - sh0;
            int idx[kTbPts];
            double d[kTbPts];
#pragma unroll
            for (int j = 0; j < kTbPts; ++j) {
                const double dx = x[j] - ax, dy = y[j] - ay, dz = z[j] - az;
                idx[j] = table_interval(tab, a, sqrt_nocall(fma(dz, dz, fma(dy, dy, dx * dx))), d[j]);
            }
            for (int k = 0; k < ns; ++k) {
                const double* c = tab.shell_coef + tab.shell_coef_off[sh0 + k];
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < kTbPts; ++j) s = fma(t[j], table_cubic(c + 4 * idx[j], d[j]), s);
                s = warp_allsum(s);
                if (lane == 0) s_acc[warp][k] = s;
            }
            __syncthreads();
            for (int k = threadIdx.x; k < ns; k += kTbThreads) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kTbWarps; ++w) tot += s_acc[w][k];
                my_partial[sh0 + k] += tot;
            }
            __syncthreads();
        }
    }
}

__global__ void table_fold_kernel(int nblk, int nout, const double* __restrict__ partial, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nout) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[int64_t(b) * nout + c];
    out[c] = s;
}

}  // namespace hp

using namespace hp;

extern "C" int32_t hp_molgrid_num_blocks(int64_t npts);

extern "C" int hp_shell_moments_table(int64_t npts, const double* px, const double* py, const double* pz,
                                      int32_t natom, const double* atom_xyz, const int32_t* atom_shell_offsets,
                                      const int32_t* knot_offsets, const double* knots, const int32_t* lut_meta,
                                      const uint16_t* lut, const int64_t* shell_coef_offsets,
                                      const double* shell_coef, const double* rho, const double* molw,
                                      const double* promol, double density_cutoff, int32_t power,
                                      int32_t nshell, int32_t nshell_max_per_atom, double* partial, double* out,
                                      void* stream) {
    HP_REQUIRE(npts > 0 && natom > 0 && nshell > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_shell_offsets && knot_offsets && knots && lut_meta && lut &&
                   shell_coef_offsets && shell_coef && rho && molw && promol && partial && out, "null input");
    HP_REQUIRE(power == 1 || power == 2, "power must be 1 or 2");
    HP_REQUIRE(nshell_max_per_atom > 0 && nshell_max_per_atom <= kTbMaxShells, "too many shells per atom (<= 64)");
    const int nblk = hp_molgrid_num_blocks(npts);
    TableArgs tab{knot_offsets, knots, lut_meta, lut, reinterpret_cast<const long long*>(shell_coef_offsets), shell_coef};
    table_moments_kernel<<<nblk, kTbThreads, 0, as_stream(stream)>>>(npts, px, py, pz, natom, atom_xyz,
                                                                     atom_shell_offsets, tab, rho, molw, promol,
                                                                     density_cutoff, power, nshell, partial);
    HP_LAUNCH_CHECK("table_moments_kernel");
    table_fold_kernel<<<(nshell + 255) / 256, 256, 0, as_stream(stream)>>>(nblk, nshell, partial, out);
    HP_LAUNCH_CHECK("table_fold_kernel");
    return HP_OK;
}
