// Cut-off ("local grid") variant of the fused promolecule / owner-weight / entropy pass.
//
// Semantics = the reference's local-grid design (commented block core/stockholder.py:45-112 +
// qc-grid Grid.get_localgrid): atom a contributes only to the points with
//     ((dx*dx + dy*dy) + dz*dz) <= radius*radius        (unfused FP64, exactly row L's rule)
// and its own weight is zero outside that ball.  With radius = inf this is the dense pass.
//
// Shell screening (optional, `shell_skip`): shell k of an atom is dropped for a whole chunk when, at
// the chunk's minimum distance to that atom, it is below 2^-nbits (default 2^-100) of the atom's most
// diffuse shell -- it could not change the FP64 value of the atom's pro-atom sum (the chance of a
// last-bit flip is 2^(53-nbits) per evaluation).  The thresholds come from hp_shell_screen.  This is
// what makes far-away core shells and tight Gaussians free.
//
// Atom screening (optional, `atom_eps` > 0, all amplitudes >= 0): atom b is dropped for a whole chunk
// when an upper bound of its pro-atom there, ub_b = sum_k A_k exp(-alpha_k x_min), is below
// atom_eps * lb, with lb a lower bound of the promolecule on the chunk: the larger of the owner
// atom's own pro-atom at the chunk's outer radius and the block minimum of the running sums after
// the previous tile (the sums only grow).  With atom_eps <= 2^-54 / natom all dropped terms together
// are below half an ulp of the final sum, i.e. far below the rounding noise a sequential FP64 sum
// of natom terms carries anyway (~sqrt(natom) ulp).  Measured on config 5 (2,000 atoms, 58.2 M
// points): 98.5 % of the promolecule values are bit-identical to the unscreened pass, the rest
// differ by <= 50 ulp (different rounding sequence), charges by 1.8e-15; 47.9 % of the pairs are
// evaluated.  lb <= 1e-80 disables the test, which keeps the +1e-100 offsets literal.
//
// Work skipping happens per block and is conservative; the per-pair test above is exact:
//   a chunk is up to 1,024 consecutive grid points of ONE owner atom (chunks never straddle atom
//   blocks), i.e. a few radial shells, r_min <= |p - R_o| <= r_max, so atom a can reach it only if
//   r_min - radius <= |R_a - R_o| <= r_max + radius.  Atoms passing this annulus test are compacted
//   IN ATOM ORDER into the shared-memory tile (warp ballots), so the sequential summation order of
//   the reference is preserved.  Evaluated pairs are counted for the benchmark's "evals" metric.
//
// Scheduling: chunks cost between a few dozen and natom atom evaluations, so blocks take them from
// a per-launch work counter (the last slot of the caller's chunk_scratch, zeroed on the launch
// stream: launches of different tables / streams never share it) instead of a fixed stride (the slowest block of a static split was ~10 %
// behind the mean on config 5).  Reductions stay bit-reproducible: every chunk writes its entropy
// term to its own slot, a second kernel folds the slots into the partial-sum buffer in a fixed
// order, and the pair counters are integers.
//
// Inner loop (one candidate atom x 4 points per thread): the atom record (centre, first shell,
// shell range, guard flag) is three LDS.128 issued from PTX into the registers of the record they
// replace, right after their last use -- no rotation moves; the exp constants sit in registers.
#include "hp_promol_common.cuh"

namespace hp {

#ifndef HP_LOC_THREADS
#define HP_LOC_THREADS 128
#endif
#ifndef HP_LOC_BLOCKS
#define HP_LOC_BLOCKS 3
#endif
// 128 threads x 3 blocks/SM (<= 170 registers per thread): 3 warps x 8 points = 24 independent FP64
// dependency chains per scheduler; with 2 blocks (16 chains) the FP64 pipe waits on its own latency
// (B200, config 5: 112.4 -> 103.2 ms per iteration, profiles/r2_hot_kernel_variants.txt)
constexpr int kLocThreads = HP_LOC_THREADS;
constexpr int kLocBlocksPerSM = HP_LOC_BLOCKS;
#ifndef HP_LOC_PTS
#define HP_LOC_PTS 8
#endif
#ifndef HP_LOC_FUSED_SUM
#define HP_LOC_FUSED_SUM 1  // dense guard-free path: shells accumulate straight into the running sum
#endif
constexpr int kLocPts = HP_LOC_PTS;
constexpr int kLocSpan = kLocThreads * kLocPts;  // points per chunk
constexpr int kLocTileAtoms = kLocThreads;       // candidate atoms per shared-memory tile (one per thread)
constexpr int kLocTileShells = 1024;
constexpr double kDist2Bias = 1e-300;  // see the distance computation of the candidate loop

// Candidate record in shared memory: 48 bytes = three 16-byte loads.
struct __align__(16) LocAtom {
    double x, y;
    double z;
    int s0;  // first kept shell in s_AB (tile-relative); sign bit set = guarded evaluation
    int ns;  // kept shells
    double A0, alpha0;  // first kept shell (0, 0 if none)
};

__device__ __forceinline__ double dist2_unfused3(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ void lds_f64x2(unsigned addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void lds_z_pack(unsigned addr, double& z, int& s0, int& ns) {
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(z) : "r"(addr));
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(s0), "=r"(ns) : "r"(addr + 8));
}

#ifndef HP_LOC_EXP2
#define HP_LOC_EXP2 0  // 1: base-2 exponential in the candidate loop (hp_math.cuh), for A/B runs
#endif
#ifndef HP_LOC_EXP_TWO_STEP
#define HP_LOC_EXP_TWO_STEP 0  // 1: two-constant Cody-Waite reduction (16 FP64 operations per shell instead of 15)
#endif

#if HP_LOC_EXP2
using LocExpConsts = Exp2Consts;
// the loop's exponent argument is y = -(alpha log2 e) * r
__device__ __forceinline__ double loc_neg_exponent(double alpha, const LocExpConsts& c) { return -alpha * c.log2e; }
template <bool GUARD>
__device__ __forceinline__ double exp_regs(double y, const LocExpConsts& c) {
    const double e = exp2_neg_poly_regs(y, c);
    if (!GUARD) return e;
    return exp2_arg_tiny(y) ? 0.0 : e;
}
#else
using LocExpConsts = ExpConsts;
__device__ __forceinline__ double loc_neg_exponent(double alpha, const LocExpConsts&) { return -alpha; }
// exp(x) for x <= 0 with the constants in registers; GUARD adds the underflow flush of exp_neg_poly.
template <bool GUARD>
__device__ __forceinline__ double exp_regs(double x, const LocExpConsts& c) {
    const double e = exp_neg_poly_regs<HP_LOC_EXP_TWO_STEP != 0>(x, c);
    if (!GUARD) return e;
    return exp_arg_tiny(x) ? 0.0 : e;
}
#endif

// the general-order functor (pow per shell) needs more registers: 2 blocks per SM there
constexpr int loc_blocks_per_sm(int functor) { return functor == HP_FUNCTOR_GENERAL ? 2 : kLocBlocksPerSM; }

template <int F, bool LOCAL>
__global__ void __launch_bounds__(kLocThreads, loc_blocks_per_sm(F))
promol_weights_local_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                            const double* __restrict__ pz, int64_t point_base, int natom,
                            const double* __restrict__ atom_xyz, const int64_t* __restrict__ atom_pt_off,
                            const int* __restrict__ atom_sh_off, const double* __restrict__ shell_A,
                            const double* __restrict__ shell_alpha, const double* __restrict__ shell_order,
                            int ntile, const int* __restrict__ tile_off, const double* __restrict__ rho,
                            const double* __restrict__ molw, double density_cutoff, double promol_offset,
                            double radius, const double* __restrict__ shell_skip, double atom_eps,
                            int atom_lo, int natom_local, const int64_t* __restrict__ chunk_off,
                            const int64_t* __restrict__ chunk_order, double* __restrict__ promol_out, double* __restrict__ w_out,
                            double* __restrict__ chunk_entropy,
                            unsigned long long* __restrict__ work_counter,
                            unsigned long long* __restrict__ pair_counters) {
    __shared__ LocAtom s_atoms[kLocTileAtoms + 1];  // +1: sentinel for the prefetch
    __shared__ double2 s_AB[kLocTileShells];
    __shared__ double s_N[(F == HP_FUNCTOR_GENERAL) ? kLocTileShells : 1];
    __shared__ double s_red[32];
    __shared__ double s_geom[5];  // owner centre x,y,z, r_min, r_max
    __shared__ long long s_chunk[3];  // chunk id, first local point, one past the last local point
    __shared__ int s_owner;
    __shared__ int s_wcnt[kLocThreads / 32];
    __shared__ int s_wsh[kLocThreads / 32];
    __shared__ int s_ncand;
    __shared__ double s_wmin[kLocThreads / 32];  // per-warp minimum of the running sums
    __shared__ double s_lb;                      // owner-based lower bound of the promolecule

    const double rc2 = radius * radius;
    const long long nchunk = chunk_off[natom_local];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned atoms_addr = static_cast<unsigned>(__cvta_generic_to_shared(s_atoms));
    const unsigned ab_addr = static_cast<unsigned>(__cvta_generic_to_shared(s_AB));
    unsigned long long pairs = 0, shells = 0;
    LocExpConsts ec;
    ec.load();
    // shell_skip[nshell_total] is the "negative amplitude seen" flag written by hp_shell_screen
    const bool may_screen_atoms = atom_eps > 0.0 && F != HP_FUNCTOR_GENERAL && shell_skip &&
                                  shell_skip[atom_sh_off[natom]] == 0.0;

    for (;;) {
        __syncthreads();  // everybody is done with the previous chunk's shared state
        if (threadIdx.x == 0) {
            const long long ticket = static_cast<long long>(atomicAdd(work_counter, 1ull));
            // expensive chunks (outer radial shells: nothing can be screened) are handed out first,
            // so that the last blocks finish on cheap ones
            const long long c = (ticket < nchunk && chunk_order) ? chunk_order[ticket] : ticket;
            s_chunk[0] = c;
            if (c < nchunk) {
                int lo = 0, hi = natom_local;  // owner: last local atom with chunk_off[a] <= c
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (chunk_off[mid] <= c) lo = mid; else hi = mid;
                }
                const int o = atom_lo + lo;
                const long long first = (atom_pt_off[o] - point_base) + (c - chunk_off[lo]) * kLocSpan;
                const long long end = atom_pt_off[o + 1] - point_base;
                s_chunk[1] = first;
                s_chunk[2] = end < first + kLocSpan ? end : first + kLocSpan;
                s_owner = o;
                s_geom[0] = atom_xyz[3 * o];
                s_geom[1] = atom_xyz[3 * o + 1];
                s_geom[2] = atom_xyz[3 * o + 2];
            }
        }
        __syncthreads();
        const long long chunk = s_chunk[0];
        if (chunk >= nchunk) break;
        const long long p_first = s_chunk[1], p_end = s_chunk[2];
        const int owner = s_owner;

        double x[kLocPts], y[kLocPts], z[kLocPts], pro[kLocPts];
        int nlive = 0;
#pragma unroll
        for (int j = 0; j < kLocPts; ++j) {
            const long long p = p_first + j * kLocThreads + threadIdx.x;
            const long long q = p < p_end ? p : (p_end - 1);
            nlive += p < p_end;
            x[j] = px[q]; y[j] = py[q]; z[j] = pz[q];
            pro[j] = 0.0;
        }
        // ---- chunk geometry: radial extent around the owner -------------------------------------
        double rmin, rmax;
        {
            double lo = 1e300, hi = 0.0;
#pragma unroll
            for (int j = 0; j < kLocPts; ++j) {
                const double d = sqrt(dist2_unfused3(x[j] - s_geom[0], y[j] - s_geom[1], z[j] - s_geom[2]));
                lo = fmin(lo, d);
                hi = fmax(hi, d);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
                hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
            }
            if (lane == 0) {
                s_red[warp] = lo;
                s_red[16 + warp] = hi;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double a = s_red[0], b = s_red[16];
                for (int w = 1; w < kLocThreads / 32; ++w) {
                    a = fmin(a, s_red[w]);
                    b = fmax(b, s_red[16 + w]);
                }
                s_geom[3] = a;
                s_geom[4] = b;
                if (may_screen_atoms) {  // the owner's pro-atom at the chunk's outer radius
                    const double xo = (F == HP_FUNCTOR_GAUSS) ? b * b : b;
                    double lb = 0.0;
                    for (int k = atom_sh_off[owner]; k < atom_sh_off[owner + 1]; ++k)
                        lb += shell_A[k] * exp(-shell_alpha[k] * xo * (1.0 + 1e-9));
                    s_lb = (LOCAL && b > radius) ? 0.0 : lb * (1.0 - 1e-9);
                }
            }
            if (lane == 0) s_wmin[warp] = 0.0;
            __syncthreads();
            rmin = s_geom[3];
            rmax = s_geom[4];
        }
        const double slack = LOCAL ? 1e-9 * (1.0 + rmax + radius) : 0.0;

        for (int t = 0; t < ntile; ++t) {
            const int a0 = tile_off[t], a1 = tile_off[t + 1];
            __syncthreads();  // previous tile fully consumed
            // ---- ordered compaction of the atoms that can matter for this chunk ------------------
            LocAtom rec;
            bool cand = false, fast = false;
            int nkeep = 0, shell_incl = 0, gs0 = 0, ns_all = 0;
            double xmin = 0.0;
            {
                const int i = threadIdx.x;
                if (i < a1 - a0) {
                    rec.x = atom_xyz[3 * (a0 + i) + 0];
                    rec.y = atom_xyz[3 * (a0 + i) + 1];
                    rec.z = atom_xyz[3 * (a0 + i) + 2];
                    gs0 = atom_sh_off[a0 + i];
                    ns_all = atom_sh_off[a0 + i + 1] - gs0;
                    const double D = sqrt(dist2_unfused3(rec.x - s_geom[0], rec.y - s_geom[1], rec.z - s_geom[2]));
                    cand = !LOCAL || ((D >= rmin - radius - slack) && (D <= rmax + radius + slack));
                    // conservative bounds of the chunk's distance to this atom
                    const double dmin = fmax(0.0, fmax(D - rmax, rmin - D) - 1e-9 * (1.0 + rmax + D));
                    const double dmax = (D + rmax) * (1.0 + 1e-9);
                    xmin = (F == HP_FUNCTOR_GAUSS) ? dmin * dmin : dmin;
                    const double xmax = (F == HP_FUNCTOR_GAUSS) ? dmax * dmax : dmax;
                    if (cand && may_screen_atoms) {
                        double lb = s_lb, run = s_wmin[0];
                        for (int w = 1; w < kLocThreads / 32; ++w) run = fmin(run, s_wmin[w]);
                        lb = fmax(lb, run);
                        if (lb > 1e-80) {
                            double ub = 0.0;
                            for (int k = 0; k < ns_all; ++k) ub += shell_A[gs0 + k] * exp(-shell_alpha[gs0 + k] * xmin);
                            cand = !(ub * (1.0 + 1e-9) < atom_eps * lb);
                        }
                    }
                    if (cand) {
                        nkeep = ns_all;
                        double amax = 0.0;  // largest exponent among the shells that are evaluated
                        if (shell_skip && F != HP_FUNCTOR_GENERAL) {
                            nkeep = 0;
                            for (int k = 0; k < ns_all; ++k) {
                                const bool keep = !(xmin > shell_skip[gs0 + k]);
                                nkeep += keep;
                                if (keep) amax = fmax(amax, shell_alpha[gs0 + k]);
                            }
                        } else if (F != HP_FUNCTOR_GENERAL) {
                            for (int k = 0; k < ns_all; ++k) amax = fmax(amax, shell_alpha[gs0 + k]);
                        }
                        // guards can go when no exponent argument can reach the underflow range (and,
                        // in cut-off mode where distances are exact, no point can sit on this nucleus)
                        if (F != HP_FUNCTOR_GENERAL) fast = (!LOCAL || xmin > 1e-100) && amax * xmax < 700.0;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, cand);
                // inclusive warp scan of the kept-shell counts (atom order)
                shell_incl = nkeep;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, shell_incl, off);
                    if (lane >= off) shell_incl += v;
                }
                if (lane == 31) s_wsh[warp] = shell_incl;
                if (lane == 0) s_wcnt[warp] = __popc(m);
                __syncthreads();
                int off = 0, shoff = 0, tot = 0;
                for (int w = 0; w < kLocThreads / 32; ++w) {
                    if (w < warp) {
                        off += s_wcnt[w];
                        shoff += s_wsh[w];
                    }
                    tot += s_wcnt[w];
                }
                if (cand) {
                    int dst = shoff + shell_incl - nkeep;
                    rec.s0 = fast ? dst : (dst | int(0x80000000));  // sign bit = guarded evaluation
                    rec.ns = nkeep;
                    rec.A0 = 0.0;
                    rec.alpha0 = 0.0;
                    const bool screen = shell_skip && F != HP_FUNCTOR_GENERAL;
                    bool first = true;
                    for (int k = 0; k < ns_all; ++k) {
                        if (screen && xmin > shell_skip[gs0 + k]) continue;
                        const double2 ab = make_double2(shell_A[gs0 + k], shell_alpha[gs0 + k]);
                        if (first) {
                            rec.A0 = ab.x;
                            rec.alpha0 = ab.y;
                            first = false;
                        }
                        s_AB[dst] = ab;
                        if (F == HP_FUNCTOR_GENERAL) s_N[dst] = shell_order[gs0 + k];
                        ++dst;
                    }
                    s_atoms[off + __popc(m & ((1u << lane) - 1u))] = rec;
                }
                if (threadIdx.x == 0) {
                    s_ncand = tot;
                    const LocAtom sentinel = {0.0, 0.0, 0.0, 0, 0, 0.0, 0.0};
                    s_atoms[tot] = sentinel;
                }
            }
            __syncthreads();
            const int ncand = __shfl_sync(0xffffffffu, s_ncand, 0);
            pairs += static_cast<unsigned long long>(ncand) * nlive;
            if (ncand) {  // kept shells of this tile = s0 + ns of its last candidate
                const int kept = (s_atoms[ncand - 1].s0 & 0x7fffffff) + s_atoms[ncand - 1].ns;
                shells += static_cast<unsigned long long>(kept) * nlive;
            }

            // the reference's `promoldens += 1e-100` per atom (core/stockholder.py:170) is a no-op in
            // FP64 once every running sum of the block is >= 1e-80 (1e-100 < 2^-54 * 1e-80) and the
            // terms are non-negative (sums only grow): decided per tile from the block minimum
            bool offset_is_void = false;
            if (HP_LOC_FUSED_SUM && !LOCAL && may_screen_atoms && t > 0) {
                double run = s_wmin[0];
                for (int w = 1; w < kLocThreads / 32; ++w) run = fmin(run, s_wmin[w]);
                offset_is_void = run >= 1e-80 && promol_offset <= 1e-100 && promol_offset >= 0.0;
            }

            // ---- evaluation: every thread, 4 points, all candidates in atom order ---------------
            double ax, ay, az, A0, al0;
            int s0, ns;
            lds_f64x2(atoms_addr, ax, ay);
            lds_z_pack(atoms_addr + 16, az, s0, ns);
            lds_f64x2(atoms_addr + 32, A0, al0);
            for (int i = 0; i < ncand; ++i) {
                const unsigned next = atoms_addr + unsigned(i + 1) * unsigned(sizeof(LocAtom));  // [ncand] = sentinel
                double d2[kLocPts], f[kLocPts];
#pragma unroll
                for (int j = 0; j < kLocPts; ++j) {
                    const double dx = x[j] - ax, dy = y[j] - ay, dz = z[j] - az;
                    // dense pass: the squared distance carries a bias of 1e-300 (the DMUL of dx*dx
                    // becomes a DFMA), absorbed without trace by any d2 above 1e-284 and turning a grid
                    // point that sits exactly on a nucleus into r = 1e-150 (exp(-alpha r) is still
                    // exactly 1): the guard-free sqrt then needs no zero-distance test
                    d2[j] = LOCAL ? dist2_unfused3(dx, dy, dz) : fma(dz, dz, fma(dy, dy, fma(dx, dx, kDist2Bias)));
                }
                lds_f64x2(next, ax, ay);  // the centre is dead from here on: fetch the next one
                const int cs0 = s0 & 0x7fffffff, cns = ns;
                const bool guarded = __any_sync(0xffffffffu, s0 < 0);  // uniform by construction; the vote tells ptxas
                if (F == HP_FUNCTOR_GENERAL) {
                    const double2 ab0 = make_double2(A0, al0);
                    eval_proatom<F, kLocPts>(d2, cs0, cns, s_AB, s_N, f, ab0);
                    lds_z_pack(next + 16, az, s0, ns);
                    lds_f64x2(next + 32, A0, al0);
                } else if (HP_LOC_FUSED_SUM && !LOCAL && !guarded) {  // block-uniform branch
                    // dense pass, guard-free: every shell goes straight into the running sum with
                    // one DFMA (no separate pro-atom value, no final add), and the +1e-100 of
                    // update_pro is skipped once it cannot change the sum any more
                    double r[kLocPts];
#pragma unroll
                    for (int j = 0; j < kLocPts; ++j) r[j] = (F == HP_FUNCTOR_GAUSS) ? d2[j] : sqrt_fast(d2[j]);
                    const double na = loc_neg_exponent(al0, ec);
#pragma unroll
                    for (int j = 0; j < kLocPts; ++j) pro[j] = fma(A0, exp_regs<false>(na * r[j], ec), pro[j]);
                    lds_z_pack(next + 16, az, s0, ns);
                    lds_f64x2(next + 32, A0, al0);
                    for (int k = 1; k < cns; ++k) {
                        double A, al;
                        lds_f64x2(ab_addr + unsigned(cs0 + k) * 16u, A, al);
                        al = loc_neg_exponent(al, ec);
#pragma unroll
                        for (int j = 0; j < kLocPts; ++j) pro[j] = fma(A, exp_regs<false>(al * r[j], ec), pro[j]);
                    }
                    if (!offset_is_void) {
#pragma unroll
                        for (int j = 0; j < kLocPts; ++j) pro[j] += promol_offset;
                    }
                    continue;
                } else if (!guarded) {  // block-uniform branch
                    double r[kLocPts];
#pragma unroll
                    for (int j = 0; j < kLocPts; ++j) r[j] = (F == HP_FUNCTOR_GAUSS) ? d2[j] : sqrt_fast(d2[j]);
                    const double na = loc_neg_exponent(al0, ec);
#pragma unroll
                    for (int j = 0; j < kLocPts; ++j) f[j] = A0 * exp_regs<false>(na * r[j], ec);
                    lds_z_pack(next + 16, az, s0, ns);
                    lds_f64x2(next + 32, A0, al0);
                    for (int k = 1; k < cns; ++k) {
                        double A, al;
                        lds_f64x2(ab_addr + unsigned(cs0 + k) * 16u, A, al);
                        al = loc_neg_exponent(al, ec);
#pragma unroll
                        for (int j = 0; j < kLocPts; ++j) f[j] = fma(A, exp_regs<false>(al * r[j], ec), f[j]);
                    }
                } else {
                    double r[kLocPts];
#pragma unroll
                    for (int j = 0; j < kLocPts; ++j) r[j] = (F == HP_FUNCTOR_GAUSS) ? d2[j] : sqrt_nocall(d2[j]);
                    const double na = loc_neg_exponent(al0, ec);
#pragma unroll
                    for (int j = 0; j < kLocPts; ++j) f[j] = A0 * exp_regs<true>(na * r[j], ec);
                    lds_z_pack(next + 16, az, s0, ns);
                    lds_f64x2(next + 32, A0, al0);
                    for (int k = 1; k < cns; ++k) {
                        double A, al;
                        lds_f64x2(ab_addr + unsigned(cs0 + k) * 16u, A, al);
                        al = loc_neg_exponent(al, ec);
#pragma unroll
                        for (int j = 0; j < kLocPts; ++j) f[j] = fma(A, exp_regs<true>(al * r[j], ec), f[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < kLocPts; ++j) {
                    if (LOCAL) pro[j] = (d2[j] <= rc2) ? (pro[j] + f[j]) + promol_offset : pro[j];
                    else pro[j] = (pro[j] + f[j]) + promol_offset;
                }
            }
            if (may_screen_atoms) {  // block minimum of the running sums, consumed by the next tile's setup
                double lo = 1e300;
#pragma unroll
                for (int j = 0; j < kLocPts; ++j) lo = fmin(lo, pro[j]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
                if (lane == 0) s_wmin[warp] = lo;
            }
        }

        // ---- epilogue: promolecule, owner weight, entropy term of this chunk -------------------------
        double entropy_acc = 0.0;
        const double ox = s_geom[0], oy = s_geom[1], oz = s_geom[2];
        const int os0 = atom_sh_off[owner], ons = atom_sh_off[owner + 1] - os0;
#pragma unroll
        for (int j = 0; j < kLocPts; ++j) {
            const long long p = p_first + j * kLocThreads + threadIdx.x;
            if (p >= p_end) continue;
            if (promol_out) promol_out[p] = pro[j];
            if (w_out) {
                const double odx = x[j] - ox, ody = y[j] - oy, odz = z[j] - oz;
                const double d2 = LOCAL ? dist2_unfused3(odx, ody, odz) : fma(odz, odz, fma(ody, ody, odx * odx));
                double w = 0.0;
                if (!LOCAL || d2 <= rc2) {
                    const double r = (F == HP_FUNCTOR_GAUSS) ? d2 : sqrt_nocall(d2);
                    double fo = 0.0;
                    for (int k = 0; k < ons; ++k) {
                        const double2 ab = make_double2(shell_A[os0 + k], shell_alpha[os0 + k]);
                        const double n = (F == HP_FUNCTOR_GENERAL) ? shell_order[os0 + k] : 1.0;
                        fo = fma(ab.x, shell_value<F>(ab, n, r), fo);
                    }
                    w = fmin(fmax(fo / pro[j], 0.0), 1.0);
                }
                w_out[p] = w;
            }
            if (chunk_entropy) {
                const double r = rho[p];
                const bool sick = (pro[j] < density_cutoff) || (r < density_cutoff);
                if (!sick) entropy_acc += molw[p] * r * log(r / pro[j]);
            }
        }
        if (chunk_entropy) {
            const double total = block_sum(entropy_acc, s_red);
            if (threadIdx.x == 0) chunk_entropy[chunk] = total;
        }
    }

    if (pair_counters) {
        // each thread tallied (candidates x its own live points): integers, so the atomics are exact
        unsigned long long v = pairs, u = shells;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, off);
            u += __shfl_xor_sync(0xffffffffu, u, off);
        }
        if (lane == 0) {
            atomicAdd(&pair_counters[0], v);
            atomicAdd(&pair_counters[kMaxPartials], u);
        }
    }
}

// entropy_partials[i] = sum over chunks c = i, i + kMaxPartials, ... (ascending) of chunk_entropy[c]
__global__ void fold_chunk_entropy_kernel(long long nchunk, const double* __restrict__ chunk_entropy,
                                          double* __restrict__ partials) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kMaxPartials) return;
    double s = 0.0;
    for (long long c = i; c < nchunk; c += kMaxPartials) s += chunk_entropy[c];
    partials[i] = s;
}

// Screening thresholds: shell_skip[k] = largest value of the radial variable (r for Slater, r^2 for
// Gaussian shells) at which shell k is still >= 2^-nbits of the atom's most diffuse shell with a
// non-zero amplitude.  +inf = never dropped, -1 = always negligible (zero amplitude).
__global__ void shell_screen_kernel(int natom, const int* __restrict__ atom_sh_off,
                                    const double* __restrict__ A, const double* __restrict__ alpha,
                                    double nbits, double* __restrict__ skip) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= natom) return;
    const int s0 = atom_sh_off[a], s1 = atom_sh_off[a + 1];
    int ref = -1;
    bool clean = true;
    bool nonneg = true;
    for (int k = s0; k < s1; ++k) {
        clean = clean && isfinite(A[k]) && isfinite(alpha[k]) && alpha[k] >= 0.0;
        nonneg = nonneg && A[k] >= 0.0;
        if (A[k] != 0.0 && (ref < 0 || alpha[k] < alpha[ref])) ref = k;
    }
    // skip[nshell_total] != 0: some amplitude is negative / not finite, atom screening must stay off
    if (!(clean && nonneg)) skip[atom_sh_off[natom]] = 1.0;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int k = s0; k < s1; ++k) {
        double t = inf;
        if (clean && ref >= 0) {
            if (A[k] == 0.0) {
                t = -1.0;
            } else if (k != ref) {
                const double lr = log(fabs(A[k] / A[ref])) + nbits * 0.6931471805599453;
                const double da = alpha[k] - alpha[ref];
                if (da > 0.0) t = lr / da;
                else if (lr < 0.0) t = -1.0;  // same exponent, amplitude below the threshold
            }
        }
        skip[k] = t;
    }
}

}  // namespace hp

using namespace hp;

extern "C" int hp_shell_screen(int32_t natom, int32_t nshell, const int32_t* atom_shell_offsets,
                               const double* shell_A, const double* shell_alpha, double nbits,
                               double* shell_skip, void* stream) {
    HP_REQUIRE(natom > 0 && nshell >= 0 && atom_shell_offsets && shell_A && shell_alpha && shell_skip,
               "bad arguments");
    HP_REQUIRE(nbits >= 60.0, "nbits must be >= 60");
    int rc = check_cuda(cudaMemsetAsync(shell_skip + nshell, 0, sizeof(double), as_stream(stream)), "memset");
    if (rc) return rc;
    shell_screen_kernel<<<(natom + 127) / 128, 128, 0, as_stream(stream)>>>(natom, atom_shell_offsets, shell_A,
                                                                            shell_alpha, nbits, shell_skip);
    HP_LAUNCH_CHECK("shell_screen_kernel");
    return HP_OK;
}

extern "C" void hp_local_tile_limits(int32_t* max_atoms_host, int32_t* max_shells_host) {
    *max_atoms_host = kLocTileAtoms;
    *max_shells_host = kLocTileShells;
}

extern "C" int32_t hp_local_chunk_points(void) { return kLocSpan; }

extern "C" int hp_promol_weights_local(int functor, int64_t npts, const double* px, const double* py,
                                       const double* pz, int64_t point_base, int32_t natom,
                                       const double* atom_xyz, const int64_t* atom_point_offsets,
                                       const int32_t* atom_shell_offsets, const double* shell_A,
                                       const double* shell_alpha, const double* shell_order,
                                       int32_t ntile, const int32_t* tile_atom_offsets,
                                       const double* rho, const double* molw, double density_cutoff,
                                       double promol_offset, double radius, const double* shell_skip,
                                       double atom_eps, int32_t atom_lo, int32_t natom_local,
                                       const int64_t* chunk_offsets, const int64_t* chunk_order,
                                       int64_t nchunk, double* chunk_scratch, double* promol,
                                       double* at_weights, double* entropy_partials,
                                       uint64_t* pair_partials, void* stream) {
    HP_REQUIRE(npts >= 0 && natom > 0 && ntile > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_point_offsets && atom_shell_offsets, "null input");
    HP_REQUIRE(shell_A && shell_alpha && tile_atom_offsets, "null shell table");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    HP_REQUIRE(!entropy_partials || (rho && molw), "entropy needs rho and molw");
    HP_REQUIRE(radius >= 0.0, "negative radius (use +inf for the dense pass)");
    HP_REQUIRE(atom_eps >= 0.0 && atom_eps < 1e-9, "atom_eps must be in [0, 1e-9)");
    HP_REQUIRE(atom_lo >= 0 && natom_local >= 0 && atom_lo + natom_local <= natom, "bad local atom range");
    HP_REQUIRE(nchunk >= 0 && (nchunk == 0 || chunk_offsets), "chunk offsets missing");
    HP_REQUIRE(nchunk == 0 || chunk_scratch, "chunk_scratch (nchunk + 1 doubles) is required");
    cudaStream_t st = as_stream(stream);
    int rc = HP_OK;
    if (pair_partials) {
        rc = check_cuda(cudaMemsetAsync(pair_partials, 0, sizeof(uint64_t) * 2 * kMaxPartials, st), "memset");
        if (rc) return rc;
    }
    if (npts == 0 || nchunk == 0) {
        if (entropy_partials)
            return check_cuda(cudaMemsetAsync(entropy_partials, 0, sizeof(double) * kMaxPartials, st), "memset");
        return HP_OK;
    }
    // the launch's own work counter: slot [nchunk] of chunk_scratch
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(chunk_scratch + nchunk);
    rc = check_cuda(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st), "memset");
    if (rc) return rc;
    int64_t grid = int64_t(sm_count()) * loc_blocks_per_sm(functor);
    if (grid > nchunk) grid = nchunk;
    const bool local = !isinf(radius);
    double* chunk_entropy = entropy_partials ? chunk_scratch : nullptr;
#define HP_LOC_ARGS                                                                                      \
    npts, px, py, pz, point_base, natom, atom_xyz, atom_point_offsets, atom_shell_offsets, shell_A,      \
        shell_alpha, shell_order, ntile, tile_atom_offsets, rho, molw, density_cutoff, promol_offset,    \
        radius, shell_skip, atom_eps, atom_lo, natom_local, chunk_offsets, chunk_order, promol,          \
        at_weights,                                                                                      \
        chunk_entropy, counter, reinterpret_cast<unsigned long long*>(pair_partials)
#define HP_LOC(F)                                                                                        \
    if (local) promol_weights_local_kernel<F, true><<<int(grid), kLocThreads, 0, st>>>(HP_LOC_ARGS);     \
    else promol_weights_local_kernel<F, false><<<int(grid), kLocThreads, 0, st>>>(HP_LOC_ARGS)
    switch (functor) {
        case HP_FUNCTOR_SLATER: HP_LOC(HP_FUNCTOR_SLATER); break;
        case HP_FUNCTOR_GAUSS: HP_LOC(HP_FUNCTOR_GAUSS); break;
        case HP_FUNCTOR_GENERAL: HP_LOC(HP_FUNCTOR_GENERAL); break;
        default: set_error("hp_promol_weights_local: unsupported functor %d", functor); return HP_ERR_ARG;
    }
#undef HP_LOC
#undef HP_LOC_ARGS
    HP_LAUNCH_CHECK("promol_weights_local_kernel");
    if (entropy_partials) {
        fold_chunk_entropy_kernel<<<kMaxPartials / 256, 256, 0, st>>>(nchunk, chunk_scratch, entropy_partials);
        HP_LAUNCH_CHECK("fold_chunk_entropy_kernel");
    }
    return HP_OK;
}

// Fold of the per-chunk entropy slots into the kMaxPartials partial sums, as hp_promol_weights_local does
// at its end: for a pass that was launched in several parts over disjoint chunk ranges (first iteration of
// a slab whose upload is still running), one fold over all slots gives the single launch's result bit for bit.
extern "C" int hp_fold_chunk_entropy(int64_t nchunk, const double* chunk_scratch, double* entropy_partials,
                                     void* stream) {
    HP_REQUIRE(nchunk >= 0 && entropy_partials && (nchunk == 0 || chunk_scratch), "bad arguments");
    fold_chunk_entropy_kernel<<<kMaxPartials / 256, 256, 0, as_stream(stream)>>>(nchunk, chunk_scratch, entropy_partials);
    HP_LAUNCH_CHECK("fold_chunk_entropy_kernel");
    return HP_OK;
}
