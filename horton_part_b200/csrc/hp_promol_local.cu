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
//   a chunk of consecutive grid points almost always lies on a few radial shells of ONE owner atom,
//   r_min <= |p - R_o| <= r_max, so atom a can reach it only if
//   r_min - radius <= |R_a - R_o| <= r_max + radius.  Atoms passing this annulus test are compacted
//   IN ATOM ORDER into the shared-memory tile (warp ballots), so the sequential summation order of
//   the reference is preserved.  Evaluated pairs are counted for the benchmark's "evals" metric.
#include "hp_promol_common.cuh"

namespace hp {

constexpr int kLocThreads = 256;
constexpr int kLocPts = 4;

__device__ __forceinline__ double dist2_unfused3(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

template <int F, bool LOCAL>
__global__ void __launch_bounds__(kLocThreads, 2)
promol_weights_local_kernel(int64_t npts, const double* __restrict__ px, const double* __restrict__ py,
                            const double* __restrict__ pz, int64_t point_base, int natom,
                            const double* __restrict__ atom_xyz, const int64_t* __restrict__ atom_pt_off,
                            const int* __restrict__ atom_sh_off, const double* __restrict__ shell_A,
                            const double* __restrict__ shell_alpha, const double* __restrict__ shell_order,
                            int ntile, const int* __restrict__ tile_off, const double* __restrict__ rho,
                            const double* __restrict__ molw, double density_cutoff, double promol_offset,
                            double radius, const double* __restrict__ shell_skip, double atom_eps,
                            double* __restrict__ promol_out, double* __restrict__ w_out,
                            double* __restrict__ entropy_partials,
                            unsigned long long* __restrict__ pair_partials) {
    __shared__ AtomRec s_atoms[kTileAtoms + 1];  // +1: sentinel for the prefetch
    __shared__ double2 s_AB[kTileShells];
    __shared__ double s_N[(F == HP_FUNCTOR_GENERAL) ? kTileShells : 1];
    __shared__ double s_red[32];
    __shared__ double s_geom[5];  // owner centre x,y,z, r_min, r_max
    __shared__ int s_flags[2];    // same-owner flag, owner index
    __shared__ int s_wcnt[kTileAtoms / 32];
    __shared__ int s_wsh[kTileAtoms / 32];
    __shared__ int s_ncand;
    __shared__ double s_wmin[kLocThreads / 32];  // per-warp minimum of the running sums
    __shared__ double s_lb;                      // owner-based lower bound of the promolecule

    const double rc2 = radius * radius;
    const int64_t span = int64_t(kLocThreads) * kLocPts;
    const int64_t nchunk = (npts + span - 1) / span;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double entropy_acc = 0.0;
    unsigned long long pairs = 0, shells = 0;

    for (int64_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        double x[kLocPts], y[kLocPts], z[kLocPts], pro[kLocPts];
        int nlive = 0;
#pragma unroll
        for (int j = 0; j < kLocPts; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kLocThreads + threadIdx.x;
            const int64_t q = p < npts ? p : (npts - 1);
            nlive += p < npts;
            x[j] = px[q]; y[j] = py[q]; z[j] = pz[q];
            pro[j] = 0.0;
        }
        // ---- chunk geometry: one owner? radial extent around it ---------------------------------
        __syncthreads();
        if (threadIdx.x == 0) {
            const int64_t first = point_base + chunk * span;
            int64_t last = point_base + chunk * span + span - 1;
            if (last > point_base + npts - 1) last = point_base + npts - 1;
            int own[2];
            const int64_t g2[2] = {first, last};
            for (int e = 0; e < 2; ++e) {
                int lo = 0, hi = natom;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (atom_pt_off[mid] <= g2[e]) lo = mid; else hi = mid;
                }
                own[e] = lo;
            }
            s_flags[0] = own[0] == own[1];
            s_flags[1] = own[0];
            s_geom[0] = atom_xyz[3 * own[0]];
            s_geom[1] = atom_xyz[3 * own[0] + 1];
            s_geom[2] = atom_xyz[3 * own[0] + 2];
        }
        __syncthreads();
        const bool same_owner = s_flags[0] != 0;
        double rmin = 0.0, rmax = 0.0;
        if (same_owner) {
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
                s_red[8 + warp] = hi;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double a = s_red[0], b = s_red[8];
                for (int w = 1; w < kLocThreads / 32; ++w) {
                    a = fmin(a, s_red[w]);
                    b = fmax(b, s_red[8 + w]);
                }
                s_geom[3] = a;
                s_geom[4] = b;
            }
            __syncthreads();
            rmin = s_geom[3];
            rmax = s_geom[4];
        }
        const double slack = LOCAL ? 1e-9 * (1.0 + rmax + radius) : 0.0;
        // shell_skip[nshell_total] is the "negative amplitude seen" flag written by hp_shell_screen
        const bool screen_atoms = atom_eps > 0.0 && same_owner && F != HP_FUNCTOR_GENERAL && shell_skip &&
                                  shell_skip[atom_sh_off[natom]] == 0.0;
        if (screen_atoms) {
            if (threadIdx.x == 0) {  // the owner's pro-atom at the chunk's outer radius
                const int o = s_flags[1];
                const double xo = (F == HP_FUNCTOR_GAUSS) ? rmax * rmax : rmax;
                double lb = 0.0;
                for (int k = atom_sh_off[o]; k < atom_sh_off[o + 1]; ++k)
                    lb += shell_A[k] * exp(-shell_alpha[k] * xo * (1.0 + 1e-9));
                s_lb = (LOCAL && rmax > radius) ? 0.0 : lb * (1.0 - 1e-9);
            }
            if (lane == 0) s_wmin[warp] = 0.0;
        }

        for (int t = 0; t < ntile; ++t) {
            const int a0 = tile_off[t], a1 = tile_off[t + 1];
            __syncthreads();  // previous tile fully consumed
            // ---- ordered compaction of the atoms that can reach this chunk -----------------------
            AtomRec rec;
            bool cand = false, fast = false;
            unsigned m = 0;
            int nkeep = 0, shell_incl = 0, gs0 = 0;
            double xmin = 0.0, xmax = 0.0;
            if (threadIdx.x < kTileAtoms) {  // warps 0..3, warp-uniform
                const int i = threadIdx.x;
                if (i < a1 - a0) {
                    rec.x = atom_xyz[3 * (a0 + i) + 0];
                    rec.y = atom_xyz[3 * (a0 + i) + 1];
                    rec.z = atom_xyz[3 * (a0 + i) + 2];
                    gs0 = atom_sh_off[a0 + i];
                    rec.ns = atom_sh_off[a0 + i + 1] - gs0;
                    cand = true;
                    if (same_owner) {
                        const double D = sqrt(dist2_unfused3(rec.x - s_geom[0], rec.y - s_geom[1], rec.z - s_geom[2]));
                        if (LOCAL) cand = (D >= rmin - radius - slack) && (D <= rmax + radius + slack);
                        // conservative lower bound of the chunk's distance to this atom
                        const double dmin = fmax(0.0, fmax(D - rmax, rmin - D) - 1e-9 * (1.0 + rmax + D));
                        xmin = (F == HP_FUNCTOR_GAUSS) ? dmin * dmin : dmin;
                        const double dmax = (D + rmax) * (1.0 + 1e-9);
                        xmax = (F == HP_FUNCTOR_GAUSS) ? dmax * dmax : dmax;
                    }
                    if (cand && screen_atoms) {
                        double lb = s_lb;  // written before the barrier at the top of this tile
                        double run = s_wmin[0];
                        for (int w = 1; w < kLocThreads / 32; ++w) run = fmin(run, s_wmin[w]);
                        lb = fmax(lb, run);
                        if (lb > 1e-80) {
                            double ub = 0.0;
                            for (int k = 0; k < rec.ns; ++k)
                                ub += shell_A[gs0 + k] * exp(-shell_alpha[gs0 + k] * xmin);
                            cand = !(ub * (1.0 + 1e-9) < atom_eps * lb);
                        }
                    }
                    if (cand) {
                        nkeep = rec.ns;
                        double amax = 0.0;  // largest exponent among the shells that are evaluated
                        if (shell_skip && F != HP_FUNCTOR_GENERAL) {
                            nkeep = 0;
                            for (int k = 0; k < rec.ns; ++k) {
                                const bool keep = !(xmin > shell_skip[gs0 + k]);
                                nkeep += keep;
                                if (keep) amax = fmax(amax, shell_alpha[gs0 + k]);
                            }
                        } else if (F != HP_FUNCTOR_GENERAL) {
                            for (int k = 0; k < rec.ns; ++k) amax = fmax(amax, shell_alpha[gs0 + k]);
                        }
                        // guards can go when no point of the chunk can sit on this nucleus and no
                        // exponent argument can reach the underflow range
                        if (same_owner && F != HP_FUNCTOR_GENERAL) fast = xmin > 1e-100 && amax * xmax < 700.0;
                    }
                }
                m = __ballot_sync(0xffffffffu, cand);
                // inclusive warp scan of the kept-shell counts (atom order)
                shell_incl = nkeep;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, shell_incl, off);
                    if (lane >= off) shell_incl += v;
                }
                if (lane == 31) s_wsh[warp] = shell_incl;
                if (lane == 0) s_wcnt[warp] = __popc(m);
            }
            __syncthreads();
            if (threadIdx.x < kTileAtoms) {
                int off = 0, shoff = 0;
                for (int w = 0; w < warp; ++w) {
                    off += s_wcnt[w];
                    shoff += s_wsh[w];
                }
                if (cand) {
                    int dst = shoff + shell_incl - nkeep;
                    rec.s0 = fast ? dst : (dst | int(0x80000000));  // sign bit = guarded evaluation
                    const bool screen = shell_skip && F != HP_FUNCTOR_GENERAL;
                    for (int k = 0; k < rec.ns; ++k) {
                        if (screen && xmin > shell_skip[gs0 + k]) continue;
                        s_AB[dst] = make_double2(shell_A[gs0 + k], shell_alpha[gs0 + k]);
                        if (F == HP_FUNCTOR_GENERAL) s_N[dst] = shell_order[gs0 + k];
                        ++dst;
                    }
                    rec.ns = nkeep;
                    s_atoms[off + __popc(m & ((1u << lane) - 1u))] = rec;
                }
                if (threadIdx.x == 0) {
                    int tot = 0;
                    for (int w = 0; w < kTileAtoms / 32; ++w) tot += s_wcnt[w];
                    s_ncand = tot;
                    AtomRec sentinel = {0.0, 0.0, 0.0, 0, 0};
                    s_atoms[tot] = sentinel;
                }
            }
            __syncthreads();
            const int ncand = s_ncand;
            pairs += static_cast<unsigned long long>(ncand) * nlive;
            if (pair_partials) {  // kept shells of this tile = s0 + ns of its last candidate
                const int kept = ncand ? (s_atoms[ncand - 1].s0 & 0x7fffffff) + s_atoms[ncand - 1].ns : 0;
                shells += static_cast<unsigned long long>(kept) * nlive;
            }

            AtomRec nxt = s_atoms[0];
            double2 nxt_ab = s_AB[nxt.s0 & 0x7fffffff];
            for (int i = 0; i < ncand; ++i) {
                const AtomRec ar = nxt;
                const double2 ab0 = nxt_ab;
                nxt = s_atoms[i + 1];  // entry [ncand] is a sentinel
                nxt_ab = s_AB[nxt.s0 & 0x7fffffff];
                double d2[kLocPts], f[kLocPts];
#pragma unroll
                for (int j = 0; j < kLocPts; ++j) {
                    const double dx = x[j] - ar.x, dy = y[j] - ar.y, dz = z[j] - ar.z;
                    d2[j] = LOCAL ? dist2_unfused3(dx, dy, dz) : fma(dz, dz, fma(dy, dy, dx * dx));
                }
                // block-uniform branch: zero-distance / underflow guards dropped where the chunk
                // geometry rules both out
                if (ar.s0 >= 0) eval_proatom<F, kLocPts, true>(d2, ar.s0, ar.ns, s_AB, s_N, f, ab0);
                else eval_proatom<F, kLocPts>(d2, ar.s0 & 0x7fffffff, ar.ns, s_AB, s_N, f, ab0);
#pragma unroll
                for (int j = 0; j < kLocPts; ++j) {
                    if (LOCAL) pro[j] = (d2[j] <= rc2) ? (pro[j] + f[j]) + promol_offset : pro[j];
                    else pro[j] = (pro[j] + f[j]) + promol_offset;
                }
            }
            if (screen_atoms) {  // block minimum of the running sums, consumed by the next tile's setup
                double lo = 1e300;
#pragma unroll
                for (int j = 0; j < kLocPts; ++j) lo = fmin(lo, pro[j]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
                if (lane == 0) s_wmin[warp] = lo;
            }
        }

#pragma unroll
        for (int j = 0; j < kLocPts; ++j) {
            const int64_t p = chunk * span + int64_t(j) * kLocThreads + threadIdx.x;
            if (p >= npts) continue;
            if (promol_out) promol_out[p] = pro[j];
            if (w_out) {
                const int64_t g = point_base + p;
                int lo = 0, hi = natom;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (atom_pt_off[mid] <= g) lo = mid; else hi = mid;
                }
                const double odx = x[j] - atom_xyz[3 * lo], ody = y[j] - atom_xyz[3 * lo + 1],
                             odz = z[j] - atom_xyz[3 * lo + 2];
                const double d2 = LOCAL ? dist2_unfused3(odx, ody, odz) : fma(odz, odz, fma(ody, ody, odx * odx));
                double w = 0.0;
                if (!LOCAL || d2 <= rc2) {
                    const int s0 = atom_sh_off[lo], ns = atom_sh_off[lo + 1] - s0;
                    const double r = (F == HP_FUNCTOR_GAUSS) ? d2 : sqrt_nocall(d2);
                    double fo = 0.0;
                    for (int k = 0; k < ns; ++k) {
                        const double2 ab = make_double2(shell_A[s0 + k], shell_alpha[s0 + k]);
                        const double n = (F == HP_FUNCTOR_GENERAL) ? shell_order[s0 + k] : 1.0;
                        fo = fma(ab.x, shell_value<F>(ab, n, r), fo);
                    }
                    w = fmin(fmax(fo / pro[j], 0.0), 1.0);
                }
                w_out[p] = w;
            }
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
            for (int i = gridDim.x + threadIdx.x; i < kMaxPartials; i += kLocThreads) entropy_partials[i] = 0.0;
    }
    if (pair_partials) {
        // each thread tallied (candidates x its own live points): the block total is the sum.
        // Layout: [0, kMaxPartials) pairs, [kMaxPartials, 2 kMaxPartials) shell evaluations.
        unsigned long long v = pairs, u = shells;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, off);
            u += __shfl_xor_sync(0xffffffffu, u, off);
        }
        __shared__ unsigned long long s_pairs[2][kLocThreads / 32];
        if (lane == 0) {
            s_pairs[0][warp] = v;
            s_pairs[1][warp] = u;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0, tot2 = 0;
            for (int w = 0; w < kLocThreads / 32; ++w) {
                tot += s_pairs[0][w];
                tot2 += s_pairs[1][w];
            }
            pair_partials[blockIdx.x] = tot;
            pair_partials[kMaxPartials + blockIdx.x] = tot2;
        }
        if (blockIdx.x == 0)
            for (int i = gridDim.x + threadIdx.x; i < kMaxPartials; i += kLocThreads) {
                pair_partials[i] = 0ull;
                pair_partials[kMaxPartials + i] = 0ull;
            }
    }
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

extern "C" int hp_promol_weights_local(int functor, int64_t npts, const double* px, const double* py,
                                       const double* pz, int64_t point_base, int32_t natom,
                                       const double* atom_xyz, const int64_t* atom_point_offsets,
                                       const int32_t* atom_shell_offsets, const double* shell_A,
                                       const double* shell_alpha, const double* shell_order,
                                       int32_t ntile, const int32_t* tile_atom_offsets,
                                       const double* rho, const double* molw, double density_cutoff,
                                       double promol_offset, double radius, const double* shell_skip,
                                       double atom_eps, double* promol, double* at_weights,
                                       double* entropy_partials,
                                       uint64_t* pair_partials, void* stream) {
    HP_REQUIRE(npts >= 0 && natom > 0 && ntile > 0, "bad sizes");
    HP_REQUIRE(px && py && pz && atom_xyz && atom_point_offsets && atom_shell_offsets, "null input");
    HP_REQUIRE(shell_A && shell_alpha && tile_atom_offsets, "null shell table");
    HP_REQUIRE(functor != HP_FUNCTOR_GENERAL || shell_order, "general functor needs shell_order");
    HP_REQUIRE(!entropy_partials || (rho && molw), "entropy needs rho and molw");
    HP_REQUIRE(radius >= 0.0, "negative radius (use +inf for the dense pass)");
    HP_REQUIRE(atom_eps >= 0.0 && atom_eps < 1e-9, "atom_eps must be in [0, 1e-9)");
    cudaStream_t st = as_stream(stream);
    if (npts == 0) {
        if (entropy_partials) {
            int rc = check_cuda(cudaMemsetAsync(entropy_partials, 0, sizeof(double) * kMaxPartials, st), "memset");
            if (rc) return rc;
        }
        if (pair_partials)
            return check_cuda(cudaMemsetAsync(pair_partials, 0, sizeof(uint64_t) * 2 * kMaxPartials, st), "memset");
        return HP_OK;
    }
    const int64_t span = int64_t(kLocThreads) * kLocPts;
    int64_t grid = (npts + span - 1) / span;
    int64_t cap = int64_t(sm_count()) * 2;
    if (cap > kMaxPartials) cap = kMaxPartials;
    if (grid > cap) grid = cap;
    const bool local = !isinf(radius);
#define HP_LOC_ARGS                                                                                      \
    npts, px, py, pz, point_base, natom, atom_xyz, atom_point_offsets, atom_shell_offsets, shell_A,      \
        shell_alpha, shell_order, ntile, tile_atom_offsets, rho, molw, density_cutoff, promol_offset,    \
        radius, shell_skip, atom_eps, promol, at_weights, entropy_partials,                              \
        reinterpret_cast<unsigned long long*>(pair_partials)
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
    return HP_OK;
}
