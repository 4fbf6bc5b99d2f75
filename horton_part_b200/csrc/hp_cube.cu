// AIM quantities on arbitrary points (the uniform grid of `part-cube`): per-atom pro-atoms rho0_a, the
// promolecule, and the atoms-in-molecules densities w_a * rho (scripts/generate_cube.py:140-157, 213-227 of
// the reference: a (natom, Npts) distance array, one basis evaluation per atom, np.sum over atoms + 1e-100,
// rho0 / promol * density).  One thread per point walks the atoms in order with the shell table staged in
// shared memory (the same (A, alpha, n) table the partitioning kernels use); the natom x Npts outputs are
// written once, coalesced along the points.  Bound: HBM (16 B written per atom x point pair, 24 B with the
// re-read of rho0 for the second output) against ~40 flop per pair and shell.
#include "hp_promol_common.cuh"

namespace hp {

constexpr int kCubeThreads = 256;

template <int F>
__global__ void __launch_bounds__(kCubeThreads)
aim_on_points_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                     const double* __restrict__ pz, int natom, const double* __restrict__ atom_xyz,
                     const int* __restrict__ atom_sh_off, const double* __restrict__ shell_A,
                     const double* __restrict__ shell_alpha, const double* __restrict__ shell_order,
                     int ntile, const int* __restrict__ tile_off, const double* __restrict__ density,
                     double promol_offset, double* __restrict__ rho0_out, double* __restrict__ promol_out,
                     double* __restrict__ aim_out) {
    __shared__ AtomRec s_atoms[kTileAtoms];
    __shared__ double2 s_AB[kTileShells];
    __shared__ double s_N[(F == HP_FUNCTOR_GENERAL) ? kTileShells : 1];
    const int64_t p = int64_t(blockIdx.x) * kCubeThreads + threadIdx.x;
    const bool live = p < npts;
    const double x = live ? px[p] : 0.0, y = live ? py[p] : 0.0, z = live ? pz[p] : 0.0;
    double pro = 0.0;
    for (int t = 0; t < ntile; ++t) {
        const int a0 = tile_off[t], a1 = tile_off[t + 1];
        const int sh0 = atom_sh_off[a0], sh1 = atom_sh_off[a1];
        __syncthreads();
        for (int i = threadIdx.x; i < a1 - a0; i += kCubeThreads) {
            AtomRec rec;
            rec.x = atom_xyz[3 * (a0 + i)];
            rec.y = atom_xyz[3 * (a0 + i) + 1];
            rec.z = atom_xyz[3 * (a0 + i) + 2];
            rec.s0 = atom_sh_off[a0 + i] - sh0;
            rec.ns = atom_sh_off[a0 + i + 1] - atom_sh_off[a0 + i];
            s_atoms[i] = rec;
        }
        for (int k = threadIdx.x; k < sh1 - sh0; k += kCubeThreads) {
            s_AB[k] = make_double2(shell_A[sh0 + k], shell_alpha[sh0 + k]);
            if (F == HP_FUNCTOR_GENERAL) s_N[k] = shell_order[sh0 + k];
        }
        __syncthreads();
        if (!live) continue;
        for (int i = 0; i < a1 - a0; ++i) {
            const AtomRec at = s_atoms[i];
            const double dx = x - at.x, dy = y - at.y, dz = z - at.z;
            // np.linalg.norm: sqrt of the plain sum of squares (no FMA contraction of the reference's sum)
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            const double r = (F == HP_FUNCTOR_GAUSS) ? d2 : sqrt(d2);
            double f = 0.0;  // compute_proatom_dens: y += shell, sequential in the shell index
            for (int k = 0; k < at.ns; ++k) {
                const double2 ab = s_AB[at.s0 + k];
                const double n = (F == HP_FUNCTOR_GENERAL) ? s_N[at.s0 + k] : 1.0;
                f += ab.x * shell_value<F>(ab, n, r);
            }
            if (rho0_out) rho0_out[int64_t(a0 + i) * npts + p] = f;
            pro += f;  // np.sum(rho0, axis=0): rows added in atom order
        }
    }
    if (!live) return;
    pro += promol_offset;
    if (promol_out) promol_out[p] = pro;
    if (aim_out) {
        const double rho = density[p];
        for (int a = 0; a < natom; ++a)  // weights_funcs = rho0 / promol; aim_rho = weights_funcs * density
            aim_out[int64_t(a) * npts + p] = (rho0_out[int64_t(a) * npts + p] / pro) * rho;
    }
}

}  // namespace hp

using namespace hp;

extern "C" int hp_aim_on_points(int functor, int64_t npts, const double* px, const double* py, const double* pz,
                                int32_t natom, const double* atom_xyz, const int32_t* atom_shell_offsets,
                                const double* shell_A, const double* shell_alpha, const double* shell_order,
                                int32_t ntile, const int32_t* tile_atom_offsets, const double* density,
                                double promol_offset, double* rho0, double* promol, double* aim_rho, void* stream) {
    HP_REQUIRE(npts >= 0 && natom > 0 && ntile > 0, "bad sizes");
    if (npts == 0) return HP_OK;
    HP_REQUIRE(px && py && pz && atom_xyz && atom_shell_offsets && shell_A && shell_alpha && tile_atom_offsets,
               "null input");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    HP_REQUIRE(!aim_rho || (rho0 && density), "aim_rho needs rho0 and density");
    HP_REQUIRE(rho0 || promol, "nothing to compute");
    const int64_t blocks = (npts + kCubeThreads - 1) / kCubeThreads;
    HP_REQUIRE(blocks < (int64_t(1) << 31), "too many points for one launch");
#define HP_CUBE(F)                                                                                              \
    aim_on_points_kernel<F><<<int(blocks), kCubeThreads, 0, as_stream(stream)>>>(                               \
        npts, px, py, pz, natom, atom_xyz, atom_shell_offsets, shell_A, shell_alpha, shell_order, ntile,        \
        tile_atom_offsets, density, promol_offset, rho0, promol, aim_rho)
    switch (functor) {
        case HP_FUNCTOR_SLATER: HP_CUBE(HP_FUNCTOR_SLATER); break;
        case HP_FUNCTOR_GAUSS: HP_CUBE(HP_FUNCTOR_GAUSS); break;
        case HP_FUNCTOR_GENERAL: HP_CUBE(HP_FUNCTOR_GENERAL); break;
        default: set_error("hp_aim_on_points: unsupported functor %d", functor); return HP_ERR_ARG;
    }
#undef HP_CUBE
    HP_LAUNCH_CHECK("aim_on_points_kernel");
    return HP_OK;
}
