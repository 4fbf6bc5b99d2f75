/*
 * hp_b200.h -- C ABI of the B200-native stockholder-iteration hot path.
 *
 * The reference (horton-part, pure Python) has no FFI: its boundary is the WPart class API
 * (SURVEY.md section 8b).  These entry points are what a ctypes binding inside the reference's
 * own classes would call to replace its NumPy passes; each one cites the reference code it
 * replaces (paths relative to /root/reference/src/horton_part).  INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - the caller (PyTorch on the Python side) owns all buffers, the library allocates nothing;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless the function's comment says so;
 *   - return value: 0 on success, non-zero on failure with a thread-local message available from
 *     hp_last_error();
 *   - all floating point data is IEEE binary64; indices are int64_t (grid points) or int32_t
 *     (atoms, shells).
 */
#ifndef HP_B200_H
#define HP_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define HP_API __attribute__((visibility("default")))
#else
#define HP_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define HP_OK 0
#define HP_ERR_ARG 1
#define HP_ERR_CUDA 2

/* radial form of a pro-atom shell  A * exp(-alpha * r^n)  */
#define HP_FUNCTOR_SLATER 1  /* n = 1 for every shell: MBIS (mbis.py:263-289), slater basis     */
#define HP_FUNCTOR_GAUSS 2   /* n = 2 for every shell: gauss basis (core/basis.py:161-171)      */
#define HP_FUNCTOR_GENERAL 3 /* per-shell real n: NLIS/GMBIS (nlis.py:303-328), custom tables   */
#define HP_FUNCTOR_SPLINE 4  /* piecewise cubic in r: ISA / Hirshfeld(-I) (core/stockholder.py:271-350) */

/* flags returned per atom by the radial solvers */
#define HP_SOLVE_NOT_CONVERGED 1u /* hit the inner iteration cap (mbis.py:162, alisa.py:290)   */
#define HP_SOLVE_POP_MISMATCH 2u  /* |sum N_k - pop| > tolerance (mbis.py:157, utils.py:434)   */
#define HP_SOLVE_NONFINITE 4u

HP_API const char* hp_last_error(void);
HP_API int hp_abi_version(void);
/* Number of SMs etc. of the current device: out_host[0]=SM count, [1]=cc major, [2]=cc minor. */
HP_API int hp_device_props(int32_t* out_host);
/* Fixed number of per-block partial sums hp_promol_weights writes (size the buffer with it). */
HP_API int32_t hp_num_partials(void);

/* ------------------------------------------------------------------------------------------
 * (row L) cut-off local grid index -- replaces Grid.get_localgrid / cKDTree.query_ball_point
 * (spec: commented block core/stockholder.py:84-112).  Inclusion is the UNFUSED test
 * ((dx*dx + dy*dy) + dz*dz) <= radius*radius, indices come out sorted ascending.
 *   points_xyz   (npts,3) row-major, as the reference stores grid.points
 *   center_host  3 doubles on the host
 *   begin,end    the centre atom's own slice of the molecular grid (grid.indices[a], [a+1])
 *   out_indices  (npts)  capacity; first *count entries valid
 *   out_overlap  (npts)  1 where begin <= index < end            (may be NULL)
 *   out_dist     (npts)  |r_p - center| for the selected points  (may be NULL)
 *   out_count    1 int64 on the device
 *   scratch      at least hp_local_index_scratch_bytes(npts) bytes
 */
HP_API size_t hp_local_index_scratch_bytes(int64_t npts);
HP_API int hp_build_local_index(const double* points_xyz, int64_t npts, const double* center_host,
                         double radius, int64_t begin, int64_t end, int64_t* out_indices,
                         uint8_t* out_overlap, double* out_dist, int64_t* out_count, void* scratch,
                         size_t scratch_bytes, void* stream);

/* (npts,3) row-major -> three contiguous coordinate arrays (the kernels' SoA layout). */
HP_API int hp_split_points(const double* points_xyz, int64_t npts, double* px, double* py, double* pz,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Shell tables: turn pro-atom parameters into (A, alpha, n) per shell.
 *   MBIS     propars [N,S]*K      -> A = N S^3/(8 pi), alpha = S, n = 1   (mbis.py:284-287)
 *   exp-basis c_k, fixed (alpha,n) -> A = c_k * norm_k                    (core/basis.py:161)
 *   NLIS     propars [N,S,n]*K    -> A = N n S^(3/n)/(4 pi Gamma(3/n))    (nlis.py:322-326)
 * inv_gamma[k] = 1/Gamma(3/n_k) is supplied by the host (n is fixed during the iterations).
 */
HP_API int hp_table_mbis(int32_t nshell, const double* propars, double* shell_A, double* shell_alpha,
                  void* stream);
HP_API int hp_table_scaled(int32_t nshell, const double* coeffs, const double* norms, double* shell_A,
                    void* stream);
HP_API int hp_table_nlis(int32_t nshell, const double* propars, const double* inv_gamma, double* shell_A,
                  double* shell_alpha, double* shell_order, void* stream);

/* ------------------------------------------------------------------------------------------
 * (rows a1-a7 + a10 entropy) fused promolecule / owner-weight pass over a slab of grid points.
 * Replaces WPart.calc_radial_distances (core/base.py:630-635), eval_proatom (mbis.py:263-289,
 * gisa.py:224-244, nlis.py:303-328), update_pro (core/stockholder.py:153-175),
 * update_at_weights (core/stockholder.py:352-384) and _compute_entropy (:145-151).
 *
 * For each local point p (global index point_base + p):
 *     promol = 0;  for a in 0..natom-1 (in order):  promol = (promol + rho0_a(p)) + promol_offset
 *                  (promol_offset = 1e-100 reproduces update_pro; 0 gives gLISA's calc_promol_dens,
 *                   glisa.py:346-348)
 *     w      = clip(rho0_owner(p) / promol, 0, 1)      owner = atom whose slice contains p
 *     entropy partial += molw*rho*ln(rho/promol) unless rho < cutoff or promol < cutoff
 * Atoms are streamed through shared memory in tiles (tile_atom_offsets: ntile+1 atom indices,
 * each tile's shells must fit the kernel's shared-memory budget: hp_tile_limits()).
 *   promol, at_weights, entropy_partials may each be NULL to skip that output.
 *   entropy_partials has hp_num_partials() entries; unused ones are written as 0.
 */
HP_API void hp_tile_limits(int32_t* max_atoms_host, int32_t* max_shells_host);
HP_API int hp_promol_weights(int functor, int64_t npts, const double* px, const double* py,
                      const double* pz, int64_t point_base, int32_t natom, const double* atom_xyz,
                      const int64_t* atom_point_offsets, const int32_t* atom_shell_offsets,
                      const double* shell_A, const double* shell_alpha, const double* shell_order,
                      int32_t ntile, const int32_t* tile_atom_offsets, const double* rho,
                      const double* molw, double density_cutoff, double promol_offset,
                      double* promol, double* at_weights, double* entropy_partials, void* stream);

/* Cut-off ("local grid") variant: atom a contributes to point p only if
 * ((dx*dx+dy*dy)+dz*dz) <= radius*radius (the bit-exact rule of hp_build_local_index / qc-grid
 * Grid.get_localgrid), and the owner's weight is 0 outside its own ball -- the semantics of the
 * reference's local-grid design (commented block core/stockholder.py:45-112).  Blocks skip atoms
 * that cannot reach their chunk of points (conservative annulus test around the owner atom) while
 * keeping the atom order.  radius = +inf gives the dense pass (fused distances, no inclusion test).
 * pair_partials (2 x hp_num_partials() uint64, may be NULL) receives per-block counts of the
 * atom x point pairs actually evaluated (first half) and of the shell evaluations (second half).
 * shell_skip (may be NULL; SLATER/GAUSS functors): per-shell thresholds from hp_shell_screen; a shell
 * is dropped for a chunk of points when the chunk's minimum distance to the atom (r, or r^2 for
 * Gaussians) exceeds its threshold, i.e. when it is below 2^-nbits of the atom's most diffuse shell
 * and cannot change the FP64 pro-atom sum.  shell_skip holds nshell + 1 doubles: the last one is a
 * flag hp_shell_screen sets to 1 when any amplitude is negative or not finite (nshell = total number
 * of shells = atom_shell_offsets[natom]).
 * atom_eps > 0 (needs shell_skip, all amplitudes >= 0, SLATER/GAUSS): atom b is dropped for a chunk
 * when an upper bound of its pro-atom there is below atom_eps times a lower bound of the promolecule
 * (the owner's pro-atom at the chunk's outer radius, or the block minimum of the running sums).
 * With atom_eps <= 2^-54/natom the dropped terms together are below half an ulp of the sum, far
 * below the rounding noise of a sequential FP64 sum of natom terms; 0 disables the test. */
HP_API int hp_shell_screen(int32_t natom, int32_t nshell, const int32_t* atom_shell_offsets,
                           const double* shell_A, const double* shell_alpha, double nbits,
                           double* shell_skip, void* stream);
/* Chunks: the local points of atom a (atom_point_offsets[a..a+1] - point_base) are cut into pieces of
 * hp_local_chunk_points() points; chunk_offsets[i] (natom_local + 1 entries, device) = number of chunks
 * of the local atoms before atom atom_lo + i, nchunk = chunk_offsets[natom_local].  Blocks take chunks
 * from the launch's own work counter (slot [nchunk] of chunk_scratch) in the order given by
 * chunk_order (a permutation of 0..nchunk-1, may be NULL = ascending; the host puts the expensive
 * chunks, the outer radial shells, first so that the launch does not end on them); chunk_scratch (nchunk
 * + 1 doubles, required; one buffer per concurrently running launch) receives the per-chunk entropy
 * terms, which are folded into entropy_partials in a fixed order, and its last slot is the work
 * counter, zeroed on `stream` by the call.  pair_partials must hold 2 x hp_num_partials() uint64 ([0] = pairs, [hp_num_partials()] =
 * shell evaluations of the launch; the rest is zeroed).  Tiles must respect hp_local_tile_limits(). */
HP_API void hp_local_tile_limits(int32_t* max_atoms_host, int32_t* max_shells_host);
HP_API int32_t hp_local_chunk_points(void);
HP_API int hp_promol_weights_local(int functor, int64_t npts, const double* px, const double* py,
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
                                   uint64_t* pair_partials, void* stream);
/* One fold of `nchunk` per-chunk entropy slots (chunk_scratch, as filled by hp_promol_weights_local) into the
 * entropy partial sums: used when one pass is launched in parts over disjoint ranges of local atoms (first
 * iteration of a slab whose second half is still uploading, core/device.py) -- the result equals the single
 * launch's bit for bit.  `_update_entropy`, core/iterstock.py:125-130. */
HP_API int hp_fold_chunk_entropy(int64_t nchunk, const double* chunk_scratch, double* entropy_partials,
                                 void* stream);

/* ------------------------------------------------------------------------------------------
 * (row a8) spherical average of w_a*rho over each radial shell of each atom's own atomic grid.
 * Replaces AtomGrid.spherical_average as called from mbis.py:176-184, gisa.py:283-289,
 * isa.py:104-110:  out[s] = sum_{j in shell s} at_weights[j]*rho[j]*atgrid_w[j] / r2w[s] / (4 pi),
 * and 0 where |shell_r[s]| < 1e-8.  shell_point_offsets are LOCAL point indices (nshell+1).
 */
HP_API int hp_shell_project(int32_t nshell, const int64_t* shell_point_offsets, const double* at_weights,
                     const double* rho, const double* atgrid_w, const double* shell_r,
                     const double* shell_r2w, double* out_sph_avg, void* stream);

/* (row a13 / f3: do_density_decomposition, core/base.py:637-659 with qc-grid
 * AtomGrid.radial_component_splines)  real-spherical-harmonic components of at_weights*rho on every
 * radial shell:  out[s*(lmax+1)^2 + lm] = sum_j atgrid_w_j f_j Y_lm(Omega_j) / shell_r2w[s]
 * (0 where |shell_r| < 1e-8), Y_lm orthonormal, HORTON-2 order (C_l0, C_l1, S_l1, ...), lmax <= 16.
 * shell_atom[s] = GLOBAL index of the atom that owns shell s (centre = atom_xyz[3*atom..]). */
HP_API int hp_shell_harmonics(int32_t nshell, int32_t lmax, const int64_t* shell_point_offsets,
                              const int32_t* shell_atom, const double* px, const double* py,
                              const double* pz, const double* atom_xyz, const double* at_weights,
                              const double* rho, const double* atgrid_w, const double* shell_r,
                              const double* shell_r2w, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * (rows a9 + a10 change) per-atom radial solves, one thread block per atom (one warp for radial grids
 * beyond 256 points / basis sets beyond 24 shells), all atoms in one launch.  nrad_max = largest number
 * of radial points of the atoms of the launch, nshell_max = largest number of shells of one atom (both choose
 * the kernel shape).
 * Radial data are concatenated over atoms: atom a owns entries rad_offsets[a]..rad_offsets[a+1]
 * of rad_r (rgrid.points), rad_w4 (4 pi r^2 w_rad) and sph_avg.
 *
 * Sharding: the launch covers `natom` atoms starting at global atom `atom_base`; rad_offsets and
 * the radial arrays are rank-local (indexed from 0), par_offsets / propars / pseudo_numbers /
 * charges / msd / niter / flags are global arrays indexed by the global atom.
 *
 * hp_mbis_radial_solve -- opt_mbis_propars (mbis.py:81-163) + charge (mbis.py:200-203) +
 * this atom's term of compute_change (core/iterstock.py:32-45):
 *     propars  in/out  [N,S]*K per atom at par_offsets[a]
 *     charges  out     pseudo_numbers[a] - sum rad_w4*sph_avg
 *     msd      out     int 4 pi r^2 (rho0_new - rho0_old)^2
 *     niter, flags out per atom
 */
HP_API int hp_mbis_radial_solve(int32_t natom, int32_t atom_base, const int32_t* rad_offsets, const double* rad_r,
                         const double* rad_w4, const double* sph_avg, const int32_t* par_offsets,
                         double* propars, const double* pseudo_numbers, double inner_threshold,
                         double density_cutoff, int32_t max_inner, int32_t nrad_max, int32_t nshell_max,
                         double* charges, double* msd, int32_t* niter, uint32_t* flags, void* stream);

/* hp_nlis_radial_solve -- opt_nlis_propars (nlis.py:99-194): shells (N, S, n) with n fixed,
 * propars [N,S,n]*K per atom; shell_offsets (natom+1, global) index inv_gamma = 1/Gamma(3/n). */
HP_API int hp_nlis_radial_solve(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                                const double* rad_r, const double* rad_w4, const double* sph_avg,
                                const int32_t* par_offsets, double* propars,
                                const int32_t* shell_offsets, const double* inv_gamma,
                                const double* pseudo_numbers, double inner_threshold,
                                double density_cutoff, int32_t max_inner, int32_t nrad_max, int32_t nshell_max,
                                double* charges, double* msd, int32_t* niter, uint32_t* flags, void* stream);

/* hp_lisa_sc_radial_solve -- aLISA `solver_sc` (alisa.py:193-291) / `solver_sc_1_iter` (:294-353)
 * with compute_quantities (utils.py:198-252): c_k <- sum_i w_i c_k g_k(r_i) rho_i / pro_i.
 *   bs_funcs   basis functions on the radial grids of the LOCAL atoms, K_a x nrad_a row-major per
 *              atom at bs_offsets[local atom] (gisa.py:91-106)
 *   single_update != 0 : exactly one update, no convergence test (sc-1-iter)
 *   nrad_max, nshell_max : largest radial grid / shell count among the local atoms */
HP_API int hp_lisa_sc_radial_solve(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                                   const double* rad_w4, const double* sph_avg,
                                   const int32_t* par_offsets, double* propars,
                                   const int64_t* bs_offsets, const double* bs_funcs,
                                   const double* pseudo_numbers, double inner_threshold,
                                   double density_cutoff, double population_cutoff,
                                   int32_t max_inner, int32_t single_update, int32_t nrad_max,
                                   int32_t nshell_max, double* charges, double* msd,
                                   int32_t* niter, uint32_t* flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * Spline pro-atoms (ISA / Hirshfeld / Hirshfeld-I).
 * hp_spline_build: SciPy `CubicSpline(x, y)` (not-a-knot, extrapolating) per atom, as built by
 * get_proatom_spline (core/stockholder.py:259-269).  knot_offsets (natom+1) index `knots` and
 * `values`; atom a owns n_a-1 segments of 4 coefficients (c0..c3, highest power first) starting at
 * coef[4*(knot_offsets[a]-a)].  clip_negative != 0 applies fix_proatom_rho (:218-219).  `work`
 * needs 2 doubles per knot.
 * hp_promol_weights_spline: like hp_promol_weights with rho0_a(p) = S_a(|r_p-R_a|) + proatom_offset
 * (eval_spline / eval_proatom, core/stockholder.py:271-350; proatom_offset = 1e-100) and promol_offset
 * added per atom (update_pro's `promoldens += 1e-100`, core/stockholder.py:170; 0 for gLISA's
 * calc_promol_dens, glisa.py:346-348).
 * hp_isa_update: propars_a = max(sph_avg_a, 1e-100), charge, change term (isa.py:102-122);
 * rad_w are the plain radial weights (rgrid.weights). */
HP_API int hp_spline_build(int32_t natom, const int32_t* knot_offsets, const double* knots,
                           const double* values, int32_t clip_negative, double* coef, double* work,
                           const int64_t* inv_offsets, const double* invT, int32_t nknot_max, void* stream);
/* The not-a-knot system matrix depends on the knots only, which never change during a partitioning.
 * hp_spline_system_inverse (HOST, n >= 4): invT_host[j * n + i] = (A^-1)[i][j], n x n doubles, computed in
 * long double.  With inv_offsets (per atom: offset of its matrix in the device pool invT; atoms with equal
 * knots share one) and nknot_max (largest knot count, <= 2048) hp_spline_build runs one block per atom:
 * right-hand side, s = A^-1 b and the coefficients in parallel.  NULL inv_offsets / invT: the serial Thomas
 * solve per atom (any knot count >= 2; `work` = 2 doubles per knot). */
HP_API int hp_spline_system_inverse(int32_t n, const double* knots_host, double* invT_host);
/* Interval index of the fused spline pass (the reference's eval_spline is scipy PPoly's binary
 * search, core/stockholder.py:271-302): the top bits of the double r (sign, exponent, 5 mantissa bits)
 * are a piecewise-linear, monotone log2 r; lut[(hi32(r) >> 15) - key0] is the interval holding the
 * lower edge of r's bin and a short forward scan finishes on searchsorted_right(x, r) - 1 clamped to
 * [0, n-2] -- PPoly's interval, end pieces extrapolating.  HOST helpers, one table per distinct knot
 * array: hp_spline_lut_size = number of bins, hp_spline_lut_fill writes them (uint16) and key0.
 * Per atom lut_meta = (key0, nbins, offset of its table in the device pool `lut`), 3 int32 each.
 * Atoms are streamed through shared memory in tiles of consecutive atoms (tile_atom_offsets, ntile + 1
 * entries) that respect hp_spline_tile_limits (atoms and knots per tile). */
HP_API int32_t hp_spline_lut_size(int32_t nknot, const double* knots_host);
HP_API int hp_spline_lut_fill(int32_t nknot, const double* knots_host, int32_t* key0_out,
                              uint16_t* lut_host);
HP_API void hp_spline_tile_limits(int32_t* max_atoms_host, int32_t* max_knots_host);
HP_API int hp_promol_weights_spline(int64_t npts, const double* px, const double* py,
                                    const double* pz, int64_t point_base, int32_t natom,
                                    const double* atom_xyz, const int64_t* atom_point_offsets,
                                    const int32_t* knot_offsets, const double* knots,
                                    const double* coef, const int32_t* lut_meta, const uint16_t* lut,
                                    int32_t ntile, const int32_t* tile_atom_offsets,
                                    double proatom_offset, double promol_offset, const double* rho,
                                    const double* molw, double density_cutoff, double* promol,
                                    double* at_weights, double* entropy_partials, void* stream);
HP_API int hp_isa_update(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                         const double* rad_r, const double* rad_w, const double* sph_avg,
                         const int32_t* par_offsets, double* propars, const double* pseudo_numbers,
                         double* charges, double* msd, void* stream);

/* hp_atom_weight_integrals_spline -- populations on the MOLECULAR grid for spline pro-atoms
 * (Hirshfeld with grid_type 2/3: do_populations, core/base.py:287-298, with the full-grid weights
 * of update_at_weights, core/stockholder.py:352-384):
 *     out[a] = sum_p molw[p] * dens[p] * clip((S_a(|r_p - R_a|) + proatom_offset) / promol[p], 0, 1)
 * over all npts points, without storing natom x Npts weight arrays.  `promol` comes from
 * hp_promol_weights_spline on the same coefficients; `partial` is scratch of
 * natom x hp_spline_integral_blocks(npts) doubles, folded per atom in a fixed order. */
HP_API int32_t hp_spline_integral_blocks(int64_t npts);
HP_API int hp_atom_weight_integrals_spline(int64_t npts, const double* px, const double* py,
                                           const double* pz, int32_t natom, const double* atom_xyz,
                                           const int32_t* knot_offsets, const double* knots,
                                           const double* coef, double proatom_offset,
                                           const double* dens, const double* molw,
                                           const double* promol, double* partial, double* out,
                                           void* stream);

/* ------------------------------------------------------------------------------------------
 * (row a12, and a9 on the molecular grid) reductions over ALL local grid points with the basis
 * functions regenerated in-kernel -- the reference's (M, Npts) `pro_shells` arrays
 * (glisa.py:335-344) are never built.  `partial` is scratch of hp_molgrid_num_blocks(npts) x nout
 * doubles; `out` receives the column sums in a fixed order.
 *   hp_shell_moments: out[m] = sum_p t(p) * shell_A[m] * exp(-alpha_m r^n_m),
 *       t = molw*rho/promol^power, 0 where rho < cutoff or promol < cutoff.  With shell_A = the
 *       normalisation of g_m: power 1 gives function_g's integrals (glisa.py:866-873) and minus
 *       the gradient (glisa.py:454-458).  (chunk of 1,024 points, shell) pairs whose bound
 *       |shell_A| exp(-alpha dmin^n) sum|t| is below 2^-80 are skipped (HP_B200_MOMENTS_SCREEN=0 disables).
 *   hp_atom_weight_integrals: out[a] = sum_p molw*rho*clip(rho0_a/promol, 0, 1) with
 *       rho0_a = sum_{m in a} shell_A[m] exp(...)  (update_at_weights(force_on_molgrid) +
 *       grid.integrate(at_weights*rho), glisa.py:269-278), without storing natom x Npts weights.
 *   hp_radial_change: msd[a] = int 4 pi r^2 (sum_k (c_new-c_old)_k g_k)^2 on the radial grids
 *       (core/iterstock.py:32-45 for the exponential-basis schemes). */
HP_API int32_t hp_molgrid_num_blocks(int64_t npts);
HP_API int hp_shell_moments(int functor, int64_t npts, const double* px, const double* py,
                            const double* pz, int32_t natom, const double* atom_xyz,
                            const int32_t* atom_shell_offsets, const double* shell_A,
                            const double* shell_alpha, const double* shell_order, int32_t ntile,
                            const int32_t* tile_atom_offsets, const double* rho, const double* molw,
                            const double* promol, double density_cutoff, int32_t power,
                            int32_t nshell, double* partial, double* out, void* stream);
HP_API int hp_atom_weight_integrals(int functor, int64_t npts, const double* px, const double* py,
                                    const double* pz, int32_t natom, const double* atom_xyz,
                                    const int32_t* atom_shell_offsets, const double* shell_A,
                                    const double* shell_alpha, const double* shell_order,
                                    int32_t ntile, const int32_t* tile_atom_offsets,
                                    const double* rho, const double* molw, const double* promol,
                                    double* partial, double* out, void* stream);
HP_API int hp_radial_change(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                            const double* rad_w4, const int32_t* par_offsets,
                            const int64_t* bs_offsets, const double* bs_funcs, const double* c_new,
                            const double* c_old, double* msd, void* stream);

/* (row a12, line search of the modified / quasi-Newton gLISA solvers: glisa.py:283-307
 * is_promol_valid / is_proatom_valid on the radial grids)  candidates = ncand x npar coefficient
 * vectors (global shell indexing); flags[j*natom + a_local] = 1 if candidate j makes atom a's
 * pro-atom < negative_cutoff somewhere, 2 if check_mono and it is not monotonically decaying,
 * else 0. */
HP_API int hp_radial_valid(int32_t natom, int32_t atom_base, const int32_t* rad_offsets,
                           const int32_t* par_offsets, const int64_t* bs_offsets,
                           const double* bs_funcs, const double* candidates, int32_t ncand,
                           int32_t npar, double negative_cutoff, int32_t check_mono, int32_t* flags,
                           void* stream);

/* (row a9, grid_type 2/3) one inner iteration of the per-atom fixed points on the MOLECULAR grid for
 * all atoms at once (mbis.py:128-152 / alisa.py:262-274 with rhoa = at_weights*moldens,
 * weights = grid.weights, r = radial_distances[a]; mbis.py:170-173, gisa.py:257-279).
 * Three parameter sets share the shell structure: `out` (outer iteration: defines
 * w_a = clip(rho0_a/promol,0,1)), `in` (current inner parameters), `prev` (previous inner
 * parameters, to recompute `oldpro`).  out (2*nshell + 2*natom doubles):
 *   [2m]   = sum_p molw t_m ratio        [2m+1] = sum_p molw t_m ratio r^n_m     (t_m = A_in exp(..))
 *   [2*nshell + 2a] = sum_p molw (oldpro_a - pro_a)^2      [.. + 2a + 1] = sum_p molw rhoa
 * Atoms with active[a] == 0 are skipped (NULL = all active).  `partial`: scratch of
 * hp_molgrid_num_blocks(npts) x (2*nshell + 2*natom) doubles.  Tiles must respect
 * hp_molgrid_update_tile_limits(). */
HP_API void hp_molgrid_update_tile_limits(int32_t* max_atoms_host, int32_t* max_shells_host);
HP_API int hp_molgrid_update_pass(int functor, int64_t npts, const double* px, const double* py,
                                  const double* pz, int32_t natom, const double* atom_xyz,
                                  const int32_t* atom_shell_offsets, const double* A_out,
                                  const double* alpha_out, const double* A_in, const double* alpha_in,
                                  const double* A_prev, const double* alpha_prev,
                                  const double* shell_order, const int32_t* active, int32_t ntile,
                                  const int32_t* tile_atom_offsets, const double* rho,
                                  const double* molw, const double* promol, double density_cutoff,
                                  int32_t nshell, double* partial, double* out, void* stream);

/* hp_hessian: H[m][n] = sum_p molw*rho*g_m*g_n/promol^2 (masked like hp_shell_moments), the dense
 * M x M gLISA Hessian of _working_matrix(nderiv=2) (glisa.py:459-470), row-major, both triangles
 * filled.  shell_atom[m] = atom of shell m; g_m = shell_norm[m]*exp(-alpha_m r^n).  `scratch` needs
 * hp_hessian_scratch_bytes(M) bytes.  The function enqueues one basis-panel kernel (on an internal side
 * stream, double-buffered) and one tile-product kernel per chunk of up to 64 x 1,280 grid points; it does not
 * synchronise the stream.  The side stream and its events exist once per device: like the reference's classes
 * (single-threaded, not re-entrant) two Hessians must not be enqueued concurrently on one device from two
 * host threads; consecutive calls on any streams are fine. */
HP_API size_t hp_hessian_scratch_bytes(int32_t M);
HP_API int hp_hessian(int functor, int64_t npts, const double* px, const double* py, const double* pz,
                      const double* atom_xyz, const int32_t* shell_atom, const double* shell_norm,
                      const double* shell_alpha, const double* shell_order, const double* rho,
                      const double* molw, const double* promol, double density_cutoff, int32_t M,
                      void* scratch, size_t scratch_bytes, double* H, void* stream);

/* Sum the entropy partials and sqrt(sum msd) in a fixed order: out[0] = change, out[1] = entropy. */
HP_API int hp_finish_iteration(int32_t npartial, const double* entropy_partials, int32_t natom,
                        const double* msd, double* out2, void* stream);

/* out1[0] = sum of n partials (fixed order): the rank-local entropy before an all-reduce. */
HP_API int hp_sum_partials(int32_t n, const double* partials, double* out1, void* stream);

/* (row a13, populations) out[s] = sum over points of segment s of w*f*(g or 1); seg_offsets has
 * nseg+1 LOCAL point indices.  Replaces grid.integrate(at_weights, dens) per atom
 * (core/base.py:259-264, 287-298, 320-326, glisa.py:269-278). */
HP_API int hp_segment_integrate(int32_t nseg, const int64_t* seg_offsets, const double* w,
                                const double* f, const double* g, double* out, void* stream);

/* (row a13, multipoles) out[a][:] = integrals over atom a's own slice of atgrid_w*dens*at_weights
 * times [Cartesian monomials l<=lmax (HORTON order) | real regular solid harmonics (C_l0, C_l1,
 * S_l1, ...) | r^n, n<=lmax] about R_a: the raw moments behind do_moments (core/base.py:329-402,
 * qc-grid Grid.moments).  out is (natom_global, ncart+npure+nrad) row-major; rows of the local
 * atoms atom_base..atom_base+natom-1 are written.  lmax <= 4. */
HP_API int hp_atom_moments(int32_t natom, int32_t atom_base, int32_t lmax, const int64_t* seg_offsets,
                           const double* px, const double* py, const double* pz,
                           const double* atgrid_w, const double* at_weights, const double* dens,
                           const double* atom_xyz, double* out, void* stream);

/* (section 8f-2) Becke fuzzy-cell weight of the OWNER atom at each local grid point, replacing qc-grid's
 * BeckeWeights.__call__ (call sites scripts/generate_density.py:102-111, becke.py:107-116).
 * inv_rab[a][b] = 1/|R_a-R_b| and aab[a][b] = size-adjustment parameter (clipped to +-0.45), both
 * natom x natom row-major, diagonal ignored; `order` = number of switching-polynomial iterations. */
HP_API int hp_becke_weights(int64_t npts, const double* px, const double* py, const double* pz,
                            int64_t point_base, int32_t natom, const double* atom_xyz,
                            const int64_t* atom_point_offsets, const double* inv_rab, const double* aab,
                            int32_t order, double* out, void* stream);

/* (section 8e) Neighbour counts behind the work-balanced sharding of the screened dense pass:
 * counts[a*nrad + c] = number of atoms b with |R_a - R_b|^2 <= radii2[(kind[a]*nkind + kind[b])*nrad + c]
 * (kind = class of atoms sharing radial grid and shells; nrad <= 128 thresholds = the chunks of an
 * atom block).  Everything on the device; counts as doubles. */
HP_API int hp_neighbor_counts(int32_t natom, const double* atom_xyz, const int32_t* kind, int32_t nkind,
                              int32_t nrad, const double* radii2, double* counts, void* stream);

/* Slab upload helpers (the boundary takes NumPy arrays = pageable host memory; the reference never
 * leaves the host, core/base.py:416-431 keeps `grid.points`, `grid.weights`, `moldens` as given).
 * hp_host_is_pinned: 1 if the host pointer is page-locked (registered with CUDA), else 0.
 * hp_host_to_device: copy `bytes` from host to device on `stream`.  Page-locked sources go out as one
 * asynchronous copy.  Pageable sources are pipelined through the caller's page-locked `staging`
 * buffer (two halves; `nthreads` host threads fill one half while the DMA engine drains the other)
 * and the call returns when the last chunk has left the staging buffer. */
HP_API int hp_host_is_pinned(const void* host_ptr);
HP_API int hp_host_to_device(void* dst_dev, const void* src_host, size_t bytes, void* staging,
                             size_t staging_bytes, int32_t nthreads, void* stream);

/* FP64 FMA throughput probe used by bench.py for the roofline denominator: runs `iters` dependent
 * DFMA chains (8 independent per thread) on a full grid; returns elapsed ms in *ms_host and the
 * flop count in *flops_host.  Synchronises the stream. */
HP_API int hp_dfma_probe(int32_t iters, double* sink, float* ms_host, double* flops_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device-resident outer loop (row a11): `while True: ...; if change < threshold or counter >= maxiter:
 * break` of do_partitioning (core/iterstock.py:171-188) as ONE CUDA-graph launch.
 *   hp_loop_begin   creates a graph with a conditional WHILE node and starts capturing `stream` (not the
 *                   legacy default stream) into the node's body; *loop_out receives an opaque handle.
 *   ...             the caller issues the launches of one iteration on `stream` (any hp_* entry point
 *                   that only enqueues work; no synchronisation, no allocation)
 *   hp_loop_stamp   (optional, also usable outside a capture) stamps[2 * (*counter + row_shift) + slot] =
 *                   %globaltimer in ns, for the per-iteration timings the reference records
 *                   (core/iterstock.py:135-145)
 *   hp_loop_commit  last launch of the body: history[c] = [state_vec (nvec) | out2[0] = change | out2[1] =
 *                   entropy] with c = *counter, stamps[2 c + 1] = now, *counter = c + 1, and the loop
 *                   continues iff not (change < threshold) and c + 1 < maxiter.  history holds maxiter rows.
 *   hp_loop_end     ends the capture and instantiates the graph; hp_loop_launch runs the whole loop;
 *   hp_loop_destroy releases it (and abandons an unfinished capture). */
HP_API int hp_loop_begin(void* stream, void** loop_out);
HP_API int hp_loop_stamp(void* loop, const int32_t* counter, int32_t row_shift, int32_t slot,
                         uint64_t* stamps, void* stream);
HP_API int hp_loop_commit(void* loop, int32_t nvec, const double* state_vec, const double* out2,
                          double threshold, int32_t maxiter, double* history, int32_t* counter,
                          uint64_t* stamps, void* stream);
HP_API int hp_loop_end(void* loop, void* stream);
HP_API int hp_loop_launch(void* loop, void* stream);
HP_API int hp_loop_destroy(void* loop, void* stream);

/* ------------------------------------------------------------------------------------------
 * (section 8e) multi-GPU plumbing: NCCL behind the C ABI, for consumers that shard the grid by atom blocks
 * without torch.distributed.  The reference has no communication layer; the exchange it would need is the
 * per-iteration propars / charges / change terms / entropy of core/iterstock.py:132-149, 171-188 -- one sum
 * all-reduce of the zero-filled state vector per iteration.  libnccl.so.2 is bound at run time (dlopen).
 *   hp_comm_unique_id   rank 0: 128-byte NCCL id into id128_host; ship it to the other ranks on the host
 *   hp_comm_init        every rank (its CUDA device current): *comm_out = communicator handle
 *   hp_comm_allreduce   in place over `count` doubles on `stream`; op_max = 0: sum, 1: max
 *   hp_comm_allgather   recv[r * count_per_rank ...] = send of rank r
 *   hp_comm_destroy     releases the communicator;  hp_comm_nccl_version: NCCL_VERSION_CODE or 0 */
HP_API int32_t hp_comm_nccl_version(void);
HP_API int hp_comm_unique_id(void* id128_host);
HP_API int hp_comm_init(int32_t world, int32_t rank, const void* id128_host, void** comm_out);
HP_API int hp_comm_allreduce(void* comm, double* buf, int64_t count, int32_t op_max, void* stream);
HP_API int hp_comm_allgather(void* comm, const double* send, double* recv, int64_t count_per_rank,
                             void* stream);
HP_API int hp_comm_destroy(void* comm);

/* ------------------------------------------------------------------------------------------
 * (section 8f-3) AIM quantities on the points of a uniform grid, the arrays behind `part-cube`
 * (scripts/generate_cube.py:140-157, 213-227): rho0 (natom x npts, atom-major) = the pro-atoms from the shell
 * table, promol = sum_a rho0_a + promol_offset (1e-100), aim_rho = rho0 / promol * density.  Same functors,
 * shell table and atom tiling (hp_tile_limits) as hp_promol_weights.  rho0 / promol / aim_rho may be NULL
 * when not wanted (aim_rho needs rho0). */
HP_API int hp_aim_on_points(int functor, int64_t npts, const double* px, const double* py, const double* pz,
                            int32_t natom, const double* atom_xyz, const int32_t* atom_shell_offsets,
                            const double* shell_A, const double* shell_alpha, const double* shell_order,
                            int32_t ntile, const int32_t* tile_atom_offsets, const double* density,
                            double promol_offset, double* rho0, double* promol, double* aim_rho, void* stream);

/* Block screening of the Hessian panel (on by default, HP_B200_HESSIAN_SCREEN=0 disables): a 64 x 64
 * quadrant of a 128 x 128 tile product of a sub-panel (1,280 points) is skipped when either of its 64-column
 * blocks stays below 2^-64 of the chunk's largest |Gu| -- a chunk of points sees only the basis functions of
 * the atoms around it.  hp_hessian_tiles_executed reports how many quadrant products the last call on
 * `scratch` ran, how many the unscreened product needs, and the points per sub-panel
 * (flop executed = quadrants x 2 x 64 x 64 x points). */
HP_API int hp_hessian_tiles_executed(int32_t M, int64_t npts, const void* scratch, int64_t* executed_out,
                                     int64_t* total_out, int32_t* points_per_tile_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * (row a12 with basis_type="numeric", core/basis.py:330-387) gLISA on TABULATED basis functions: shell m is a
 * piecewise cubic on its atom's knots (knot_offsets / knots as in hp_spline_build, interval tables lut_meta / lut
 * as in hp_promol_weights_spline), 4 x (n_a - 1) PPoly coefficients (highest power first) at
 * shell_coef + shell_coef_offsets[m]; atoms of one element share blocks.
 *   hp_shell_moments_table  out[m] = sum_p t(p) S_m(r_pm), t as in hp_shell_moments (function_g, gradient);
 *                           partial = hp_molgrid_num_blocks(npts) x nshell doubles of scratch
 *   hp_hessian_table        H_mn = sum_p u(p) S_m S_n, u as in hp_hessian; same scratch and tile product */
HP_API int hp_shell_moments_table(int64_t npts, const double* px, const double* py, const double* pz,
                                  int32_t natom, const double* atom_xyz, const int32_t* atom_shell_offsets,
                                  const int32_t* knot_offsets, const double* knots, const int32_t* lut_meta,
                                  const uint16_t* lut, const int64_t* shell_coef_offsets,
                                  const double* shell_coef, const double* rho, const double* molw,
                                  const double* promol, double density_cutoff, int32_t power, int32_t nshell,
                                  int32_t nshell_max_per_atom, double* partial, double* out, void* stream);
HP_API int hp_hessian_table(int64_t npts, const double* px, const double* py, const double* pz,
                            const double* atom_xyz, const int32_t* shell_atom, const int32_t* knot_offsets,
                            const double* knots, const int32_t* lut_meta, const uint16_t* lut,
                            const int64_t* shell_coef_offsets, const double* shell_coef, const double* rho,
                            const double* molw, const double* promol, double density_cutoff, int32_t M,
                            void* scratch, size_t scratch_bytes, double* H, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HP_B200_H */
