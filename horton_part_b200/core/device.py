"""Device-resident data of one partitioning job: the grid slab owned by this rank, the pro-atom
shell table, and thin wrappers that launch the C-ABI kernels on torch-owned buffers.

PyTorch is plumbing here: it allocates FP64 device buffers, provides the CUDA stream and (for
multi-GPU runs) ``torch.distributed``; every arithmetic pass over grid points is a kernel from
``libhp_b200.so``.  There is no CPU path: constructing a slab without CUDA raises.

Layout in HBM (per rank, ``n`` = local grid points, all FP64 unless noted):
    px, py, pz      n each     point coordinates, structure-of-arrays (coalesced 8-byte loads)
    molw            n          molecular quadrature weights (atomic weight x Becke weight)
    atw             n          the owner atom's un-Becke'd atomic-grid weights (grid_type=1 only)
    rho             n          molecular density
    promol, at_w    n each     outputs of the fused pass (promolecule, owner weight)
    shell tables    ~KBs       atoms' coordinates, (A, alpha, n) per shell, radial grids
Config 5 (2,000 atoms, 58.2 M points): 8 x 8 B x 58.2 M = 3.7 GB on one GPU, 0.47 GB per GPU on 8.
"""

from __future__ import annotations

import os

import numpy as np

from .. import _lib

__all__ = ["Shard", "GridSlab", "ShellTable", "require_cuda", "to_device", "stream_ptr", "estimate_dense_work"]


def require_cuda(device=None):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "horton_part_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )
    _lib.lib()  # fail loudly if the extension is not built
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def to_device(array, device, dtype=None):
    import torch

    a = np.ascontiguousarray(array, dtype=dtype)
    return torch.from_numpy(a).to(device)


def stream_ptr(device):
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def _reach(A, alpha, targets, gaussian):
    """Distances d at which sum_k A_k exp(-alpha_k x(d)) has fallen to each of ``targets`` (x = d or
    d^2): vectorised bisection on the monotone pro-atom bound; 0 where it is below the target at 0."""
    A, alpha = np.abs(np.asarray(A, float))[None, :], np.asarray(alpha, float)[None, :]
    t = np.asarray(targets, float)
    f = lambda d: np.sum(A * np.exp(-alpha * ((d * d) if gaussian else d)[:, None]), axis=1)  # noqa: E731
    lo, hi = np.zeros_like(t), np.ones_like(t)
    for _ in range(40):  # bracket
        grow = f(hi) > t
        if not grow.any():
            break
        lo = np.where(grow, hi, lo)
        hi = np.where(grow, 2.0 * hi, hi)
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        above = f(mid) > t
        lo, hi = np.where(above, mid, lo), np.where(above, hi, mid)
    return np.where(f(np.zeros_like(t)) <= t, 0.0, hi)


def estimate_dense_work(coordinates, grid, shells, gaussian=False, chunk_points=1024, eps=None, kinds=None,
                        device=None):
    """Atom x point pairs the screened dense pass will evaluate in each atom's block: the
    load-balancing weight of the sharding (geometry only, a few milliseconds).

    For every chunk of an atom's points (``chunk_points`` consecutive points = a few radial shells,
    outer radius r_c) the kernel keeps atom b if an upper bound of its pro-atom at distance
    D_ab - r_c is above ``eps`` times the owner's pro-atom at r_c (hp_promol_local.cu).  Per kind of
    atom (same radial grid and shells) that is a neighbour count inside one radius per chunk and
    neighbour kind; the radii are solved on the host, the counting runs on the device
    (``hp_neighbor_counts``).  ``shells[a] = (A_k, alpha_k)``; ``kinds[a]`` labels atoms with identical
    grids and shells (default: derived from both).  Returns None when an atom block has more than
    128 chunks (the caller then balances by points).
    """
    import torch

    xyz = np.ascontiguousarray(coordinates, dtype=np.float64)
    natom = len(xyz)
    if eps is None:
        eps = 2.0 ** -(55 + int(np.ceil(np.log2(max(natom, 2)))))
    if kinds is None:
        kinds = []
        for a in range(natom):
            g = grid.atgrids[a]
            kinds.append((id(g.rgrid), int(g.size), tuple(shells[a][0]), tuple(shells[a][1])))
    labels = {}
    kind_of = np.array([labels.setdefault(k, len(labels)) for k in kinds], dtype=np.int32)
    nkind = len(labels)
    first = [int(np.flatnonzero(kind_of == k)[0]) for k in range(nkind)]
    plans = []
    for a0 in first:
        atgrid = grid.atgrids[a0]
        idx = np.asarray(atgrid.indices, dtype=np.int64)
        r_of_point = np.repeat(np.asarray(atgrid.rgrid.points, float), np.diff(idx))
        starts = np.arange(0, int(idx[-1]), chunk_points)
        sizes = np.minimum(starts + chunk_points, int(idx[-1])) - starts
        rc = np.maximum.reduceat(r_of_point, starts)  # outer radius of every chunk
        x = rc * rc if gaussian else rc
        lb = np.sum(shells[a0][0][None, :] * np.exp(-shells[a0][1][None, :] * x[:, None]), axis=1)
        plans.append((rc, lb, sizes.astype(float)))
    nrad = max(len(p[0]) for p in plans)
    if nrad > 128:
        return None
    radii2 = np.full((nkind, nkind, nrad), -1.0)  # -1: no such chunk (never matched)
    for k, (rc, lb, _) in enumerate(plans):
        live = lb > 1e-80
        for k2, b0 in enumerate(first):
            reach = _reach(shells[b0][0], shells[b0][1], eps * np.where(live, lb, 1.0), gaussian)
            radii2[k, k2, : len(rc)] = np.where(live, (rc + reach) ** 2, np.inf)
    dev = require_cuda(device)
    counts = torch.zeros((natom, nrad), dtype=torch.float64, device=dev)
    _lib.call("hp_neighbor_counts", natom, to_device(xyz, dev), to_device(kind_of, dev, np.int32), nkind, nrad,
              to_device(radii2, dev), counts, stream_ptr(dev))  # fmt: skip
    counts = counts.cpu().numpy()
    setup_points = 55.0  # cost of screening one atom for one chunk, in point evaluations (tools/calibrate_work_model.py)
    work = np.zeros(natom)
    for k, (rc, _, sizes) in enumerate(plans):
        own = kind_of == k
        work[own] = counts[own, : len(rc)] @ sizes + setup_points * natom * len(rc)
    return work


class Shard:
    """Contiguous range of atom blocks [atom_lo, atom_hi) owned by one rank.

    The molecular grid is the concatenation of per-atom grids, so whole atom blocks are assigned to
    ranks.  Balanced by point count, or -- when ``work`` gives an estimate of the
    pairs each atom's block will evaluate (``estimate_dense_work``: the screened dense pass does less
    work for atoms at the surface of a cluster) -- by that estimate.
    """

    def __init__(self, natom, atom_point_offsets, rank=0, world=1, work=None):
        self.rank, self.world, self.natom = rank, world, natom
        off = np.asarray(atom_point_offsets, dtype=np.int64)
        if work is not None and world > 1:
            cum = np.concatenate([[0.0], np.cumsum(np.asarray(work, float))])
        else:
            cum = off.astype(float)
        total = float(cum[-1])
        # boundary b_k = first atom whose start offset >= k/world of the points (or of the work)
        targets = (np.arange(1, world) * total) / world
        cuts = np.searchsorted(cum[:-1], targets, side="left") if world > 1 else np.array([], int)
        bounds = np.concatenate([[0], cuts, [natom]]).astype(np.int64)
        bounds = np.maximum.accumulate(bounds)
        self.bounds = bounds
        self.atom_lo, self.atom_hi = int(bounds[rank]), int(bounds[rank + 1])
        self.point_lo, self.point_hi = int(off[self.atom_lo]), int(off[self.atom_hi])
        self.counts = np.diff(bounds)  # atoms per rank

    @property
    def nlocal(self):
        return self.atom_hi - self.atom_lo


class GridSlab:
    """This rank's slice of the molecular grid on the device, plus its shell / radial bookkeeping."""

    def __init__(self, grid, moldens, coordinates, device=None, shard=None, need_atgrids=True,
                 keep_aos=False):  # fmt: skip
        import torch

        self.device = require_cuda(device)
        dev = self.device
        self.natom = len(coordinates)
        self.atom_point_offsets_host = np.ascontiguousarray(grid.indices, dtype=np.int64)
        self.npts_global = int(self.atom_point_offsets_host[-1])
        self.shard = shard or Shard(self.natom, self.atom_point_offsets_host)
        lo, hi = self.shard.point_lo, self.shard.point_hi
        self.point_base = lo
        self.npts = hi - lo

        pts = np.asarray(grid.points)
        if pts.shape != (self.npts_global, 3):
            raise ValueError("grid.points must have shape (Npts, 3)")
        self.bytes_h2d = 0

        from .hostmem import upload, upload_into

        def up(a, dtype=np.float64):
            t = upload(a, dev, dtype)  # pipelined through page-locked staging when `a` is pageable
            self.bytes_h2d += t.numel() * t.element_size()
            return t

        # Large slabs go up in two parts: the points of the first local atoms now, the rest when the first
        # pass over the grid has been launched on them (ShellTable.promol_weights), so that most of the upload
        # runs under the first iteration's kernel.  Anything else that touches the point arrays first
        # completes the upload on the spot (the properties below).
        self._pending_upload = None
        self._split_atoms = self._choose_split(grid, moldens, need_atgrids)
        cut = self.npts if not self._split_atoms else int(
            self.atom_point_offsets_host[self.shard.atom_lo + self._split_atoms] - lo)
        later = []

        def up_big(host, local=False):
            """Device tensor for host[lo:hi] (host itself when ``local``); rows [0, cut) are uploaded now,
            the rest is queued for complete_upload()."""
            host = np.asarray(host)
            part = host if local else host[lo:hi]
            if part.dtype != np.float64 or not part.flags.c_contiguous:
                part = np.ascontiguousarray(part, dtype=np.float64)
            if cut == self.npts and part.nbytes < (8 << 20):
                out = torch.from_numpy(part).to(dev)
                self.bytes_h2d += out.numel() * out.element_size()
                return out
            out = torch.empty(part.shape, dtype=torch.float64, device=dev)
            self.bytes_h2d += out.numel() * out.element_size()
            if cut:
                upload_into(out[:cut], part[:cut])
            if cut < self.npts:
                later.append((out[cut:], part[cut:]))
            return out

        aos = up_big(pts)
        self._px = torch.empty(self.npts, dtype=torch.float64, device=dev)
        self._py = torch.empty_like(self._px)
        self._pz = torch.empty_like(self._px)
        if cut:
            _lib.call("hp_split_points", aos, cut, self._px, self._py, self._pz, stream_ptr(dev))
        self._aos = aos if (keep_aos or cut < self.npts) else None
        self._keep_aos = keep_aos
        self._molw = up_big(grid.weights)
        self._rho = up_big(moldens)
        self.atom_xyz = up(coordinates)
        self.atom_point_offsets = up(self.atom_point_offsets_host, np.int64)

        self._atw = None
        if need_atgrids:
            atgrids = grid.atgrids
            if atgrids is None:
                raise ValueError(
                    "Atomic grids are discarded from molecular grid object, "
                    "but are needed for local integrations."
                )
            a_lo, a_hi = self.shard.atom_lo, self.shard.atom_hi
            whole = getattr(grid, "atweights", None)
            if whole is not None and len(whole) == self.npts_global:
                self._atw = up_big(whole)
            else:
                atw = np.concatenate([atgrids[a].weights for a in range(a_lo, a_hi)]) if a_hi > a_lo else np.zeros(0)
                self._atw = up_big(atw, local=True)
            shell_parts, rad_off = [np.zeros(1, dtype=np.int64)], [0]
            rad_r, rad_w, rad_w4, rad_r2w = [], [], [], []
            per_rgrid = {}  # atoms usually share one radial grid object: its tables are formed once
            for a in range(a_lo, a_hi):
                g = atgrids[a]
                tabs = per_rgrid.get(id(g.rgrid))
                if tabs is None:
                    r, w = np.asarray(g.rgrid.points, float), np.asarray(g.rgrid.weights, float)
                    # 4 pi r^2 w: mbis.py:182, gisa.py:298;  r^2 w: qc-grid integrate_angular_coordinates
                    tabs = per_rgrid[id(g.rgrid)] = (r, w, 4 * np.pi * r**2 * w, r**2 * w)
                r, w, w4, r2w = tabs
                idx = np.asarray(g.indices, dtype=np.int64)
                shell_parts.append((self.atom_point_offsets_host[a] - lo) + idx[1:])
                rad_off.append(rad_off[-1] + len(r))
                rad_r.append(r)
                rad_w.append(w)
                rad_w4.append(w4)
                rad_r2w.append(r2w)
            cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0)  # noqa: E731
            self.rad_offsets_host = np.asarray(rad_off, dtype=np.int32)
            self.nrad_max = int(np.max(np.diff(self.rad_offsets_host), initial=1))
            self.rad_r_host, self.rad_w_host, self.rad_w4_host = cat(rad_r), cat(rad_w), cat(rad_w4)
            shell_off = np.concatenate(shell_parts)
            self.nshell = len(shell_off) - 1
            self.shell_point_offsets = up(shell_off, np.int64)
            self.rad_offsets = up(self.rad_offsets_host, np.int32)
            self.rad_r = up(self.rad_r_host)
            self.rad_w4 = up(self.rad_w4_host)
            self.rad_r2w = up(cat(rad_r2w))
            if self.nshell and int(np.max(np.diff(self.rad_offsets_host))) > 4096:
                raise ValueError("radial grids with more than 4096 points are not supported")
            self.sph_avg = torch.zeros(self.nshell, dtype=torch.float64, device=dev)

        self.promol = torch.empty(self.npts, dtype=torch.float64, device=dev)
        self.at_w = torch.empty(self.npts, dtype=torch.float64, device=dev)
        self.npartial = int(_lib.call("hp_num_partials"))
        self.entropy_partials = torch.zeros(self.npartial, dtype=torch.float64, device=dev)
        from .hostmem import drain

        if later:
            self._pending_upload = later
            # the copy stream starts after the allocations and the first part's copies queued so far -- and
            # NOT after the kernels launched later on the first part (complete_upload does not wait again)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
        else:
            drain(dev)  # asynchronous uploads from page-locked caller arrays have landed; sources released

    # -- two-part upload ------------------------------------------------------------------------------
    #: slabs below this many bytes of point data go up in one piece
    split_upload_min_bytes = 512 << 20

    def _choose_split(self, grid, moldens, need_atgrids):
        """Number of local atoms whose points are uploaded before the first pass starts (0 = no split).
        With upload time u and pass time c the best cut is the fraction u / (u + c) of the work; the pass
        costs about twice the upload from page-locked arrays and about as much as the upload from pageable
        ones, so a third / a half of the points go first."""
        from .hostmem import is_pinned

        sh = self.shard
        if (not need_atgrids or sh.nlocal < 8 or os.environ.get("HP_B200_SPLIT_UPLOAD", "1") == "0"
                or 48 * self.npts < self.split_upload_min_bytes):  # fmt: skip
            return 0
        pts = np.asarray(grid.points)
        frac = 1.0 / 3.0 if (pts.dtype == np.float64 and is_pinned(pts)) else 0.5
        off = self.atom_point_offsets_host[sh.atom_lo : sh.atom_hi + 1] - self.point_base
        m = int(np.searchsorted(off, frac * self.npts))
        return min(max(m, 1), sh.nlocal - 1)

    def complete_upload(self):
        """Upload the second part of a split slab (copy stream; the current stream waits for it)."""
        import torch

        later = self._pending_upload
        if later is None:
            return
        self._pending_upload = None
        from .hostmem import upload_into

        dev = self.device
        cur = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        for out, host in later:
            upload_into(out, host, stream=cs.cuda_stream)
        cut = self.npts - later[0][0].shape[0]
        _lib.call("hp_split_points", self._aos[cut:], self.npts - cut, self._px[cut:], self._py[cut:], self._pz[cut:],
                  cs.cuda_stream)  # fmt: skip
        cur.wait_stream(cs)
        for out, _ in later:
            out.record_stream(cs)
        if not self._keep_aos:
            self._aos.record_stream(cs)
            self._aos = None

    def finish_upload(self):
        """complete_upload + release of the page-locked sources (synchronises the current stream)."""
        from .hostmem import drain

        self.complete_upload()
        drain(self.device)

    def _whole(self, name):
        if self._pending_upload is not None:
            self.finish_upload()  # a consumer that does not know about the split: no overlap, same result
        return getattr(self, name)

    px = property(lambda self: self._whole("_px"))
    py = property(lambda self: self._whole("_py"))
    pz = property(lambda self: self._whole("_pz"))
    molw = property(lambda self: self._whole("_molw"))
    rho = property(lambda self: self._whole("_rho"))
    atw = property(lambda self: self._whole("_atw"))
    points_aos = property(lambda self: self._whole("_aos") if self._keep_aos else None)

    # ------------------------------------------------------------------------------------------
    def shell_project(self):
        """Spherical averages of at_w*rho over this rank's atoms' radial shells -> self.sph_avg."""
        _lib.call("hp_shell_project", self.nshell, self.shell_point_offsets, self.at_w, self.rho,
                  self.atw, self.rad_r, self.rad_r2w, self.sph_avg, stream_ptr(self.device))  # fmt: skip
        return self.sph_avg


class ShellTable:
    """(A, alpha, n) per shell for ALL atoms (replicated on every rank) and the atom tiling the
    promolecule kernel streams through shared memory."""

    def __init__(self, slab: GridSlab, functor: int, shells_per_atom):
        import torch

        self.slab, self.functor = slab, int(functor)
        dev = slab.device
        counts = np.asarray(shells_per_atom, dtype=np.int64)
        if len(counts) != slab.natom:
            raise ValueError("need one shell count per atom")
        self.offsets_host = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.nshell = int(self.offsets_host[-1])
        max_atoms = np.zeros(1, np.int32)
        max_shells = np.zeros(1, np.int32)
        _lib.call("hp_tile_limits", max_atoms, max_shells)
        if counts.max(initial=0) > int(max_shells[0]):
            raise ValueError("an atom has more shells than one shared-memory tile can hold")
        self._counts = counts
        self.ntile, self.tiles = self.make_tiles(int(max_atoms[0]), int(max_shells[0]))
        self.offsets = to_device(self.offsets_host, dev)
        self.A = torch.zeros(max(self.nshell, 1), dtype=torch.float64, device=dev)
        self.alpha = torch.zeros_like(self.A)
        self.order = torch.ones_like(self.A) if functor == 3 else None

    def make_tiles(self, max_atoms, max_shells):
        """Greedy split of the atom list into shared-memory tiles: (ntile, device int32 offsets)."""
        counts, natom = self._counts, self.slab.natom
        if counts.max(initial=0) > max_shells:
            raise ValueError("an atom has more shells than one shared-memory tile can hold")
        tiles, a = [0], 0
        while a < natom:
            b, nsh = a, 0
            while b < natom and b - a < max_atoms and nsh + counts[b] <= max_shells:
                nsh += counts[b]
                b += 1
            tiles.append(b)
            a = b
        return len(tiles) - 1, to_device(np.asarray(tiles, dtype=np.int32), self.slab.device)

    #: cut-off radius of the local-grid mode (None = dense, the reference's semantics)
    local_radius = None
    pair_partials = None
    #: shell screening: a shell is dropped for a chunk of points when it is below 2^-screen_bits of
    #: the atom's most diffuse shell there (cannot change the FP64 sum); None = evaluate every shell
    #: with the plain dense kernel.  Env HP_B200_SCREEN_BITS overrides (0 disables).
    screen_bits = 100.0
    skip = None
    #: atom screening (needs shell screening on): an atom is dropped for a chunk of points when an
    #: upper bound of its pro-atom there is below 2^-atom_screen_bits of a lower bound of the
    #: promolecule.  Default 2^-(55 + log2 natom): all dropped terms together stay below half an ulp
    #: of the sum (measured: <= 50 ulp = rounding-sequence noise on 1.5 % of the points of config 5,
    #: charges 2e-15).  Env HP_B200_ATOM_SCREEN=0 disables.
    atom_screen = True

    def pairs_evaluated(self):
        """atom x point pairs evaluated by the last cut-off launch on this rank."""
        return int(self.pair_partials[: self.slab.npartial].sum().item()) if self.pair_partials is not None else None

    def shells_evaluated(self):
        """shell x point evaluations of the last launch of the chunk-geometry kernel on this rank."""
        return int(self.pair_partials[self.slab.npartial :].sum().item()) if self.pair_partials is not None else None

    def promol_weights(self, density_cutoff, want_promol=True, want_weights=True, want_entropy=True,
                       promol_offset=1e-100):
        """Launch the fused promolecule / owner-weight / entropy pass over the local slab."""
        s = self.slab
        bits = self.screen_bits
        env = os.environ.get("HP_B200_SCREEN_BITS")
        if env is not None:
            bits = float(env) or None
        if self.functor == 3:
            bits = None  # mixed orders: no common radial variable to screen on
        if self.local_radius is not None or bits:
            import torch

            if self.pair_partials is None:
                self.pair_partials = torch.zeros(2 * s.npartial, dtype=torch.int64, device=s.device)
                lim_a, lim_s = np.zeros(1, np.int32), np.zeros(1, np.int32)
                _lib.call("hp_local_tile_limits", lim_a, lim_s)
                self._loc_ntile, self._loc_tiles = self.make_tiles(int(lim_a[0]), int(lim_s[0]))
                # chunks never straddle atom blocks: per-atom chunk counts of this rank's atoms
                span = int(_lib.call("hp_local_chunk_points"))
                sh = s.shard
                npt = np.diff(s.atom_point_offsets_host[sh.atom_lo : sh.atom_hi + 1])
                coff = np.concatenate([[0], np.cumsum((npt + span - 1) // span)]).astype(np.int64)
                self._loc_nchunk = int(coff[-1])
                self._loc_chunk_off = to_device(coff, s.device, np.int64)
                # hand-out order: outermost chunks of every atom first (they see all atoms, nothing
                # can be screened), the cheap inner ones last, so the launch does not end on a long chunk
                counts = np.diff(coff)
                sub = np.arange(self._loc_nchunk) - np.repeat(coff[:-1], counts)
                from_end = np.repeat(counts, counts) - 1 - sub
                order = np.argsort(from_end, kind="stable")
                self._loc_chunk_order = to_device(order, s.device, np.int64)
                self._loc_scratch = torch.zeros(self._loc_nchunk + 1, dtype=torch.float64, device=s.device)
                self._loc_split = None
                if s._pending_upload is not None:
                    # first pass of a slab that is still uploading: the atoms whose points are there / the rest
                    m = s._split_atoms
                    n_a = int(coff[m])
                    self._loc_split = (
                        m, n_a, to_device(order[order < n_a], s.device, np.int64),
                        to_device(order[order >= n_a] - n_a, s.device, np.int64),
                        to_device(coff[m:] - n_a, s.device, np.int64),
                        torch.zeros_like(self.pair_partials),
                    )
            atom_eps = 0.0
            if bits:
                if self.skip is None:
                    self.skip = torch.empty(self.nshell + 1, dtype=torch.float64, device=s.device)
                _lib.call("hp_shell_screen", s.natom, self.nshell, self.offsets, self.A, self.alpha, float(bits),
                          self.skip, stream_ptr(s.device))  # fmt: skip
                if self.atom_screen and os.environ.get("HP_B200_ATOM_SCREEN", "1") != "0":
                    atom_eps = 2.0 ** -(float(os.environ.get("HP_B200_ATOM_BITS", 55)) + int(np.ceil(np.log2(max(s.natom, 2)))))
            radius = float("inf") if self.local_radius is None else float(self.local_radius)

            def launch(atom_lo, nlocal, chunk_off, chunk_order, nchunk, scratch, pairs):
                _lib.call(
                    "hp_promol_weights_local", self.functor, s.npts, s._px, s._py, s._pz, s.point_base, s.natom,
                    s.atom_xyz, s.atom_point_offsets, self.offsets, self.A, self.alpha, self.order,
                    self._loc_ntile, self._loc_tiles, s._rho, s._molw, float(density_cutoff), float(promol_offset),
                    radius, self.skip if bits else None, atom_eps, atom_lo, nlocal,
                    chunk_off, chunk_order, nchunk, scratch,
                    s.promol if want_promol else None,
                    s.at_w if want_weights else None, s.entropy_partials if want_entropy else None,
                    pairs, stream_ptr(s.device),
                )  # fmt: skip

            if s._pending_upload is not None and getattr(self, "_loc_split", None) is not None:
                # the slab's second part is not on the device yet: pass over the atoms that are, start the
                # rest of the upload on the copy stream under that kernel, then pass over the other atoms.
                # Same chunks, same per-chunk entropy slots, one fold over all of them: results are
                # bit-identical to the single launch.
                m, n_a, order_a, order_b, coff_b, pairs_a = self._loc_split
                sh = s.shard
                launch(sh.atom_lo, m, self._loc_chunk_off, order_a, n_a, self._loc_scratch, pairs_a)
                s.complete_upload()
                launch(sh.atom_lo + m, sh.nlocal - m, coff_b, order_b, self._loc_nchunk - n_a,
                       self._loc_scratch[n_a:], self.pair_partials)
                self.pair_partials += pairs_a
                if want_entropy:
                    _lib.call("hp_fold_chunk_entropy", self._loc_nchunk, self._loc_scratch, s.entropy_partials,
                              stream_ptr(s.device))  # fmt: skip
                s.finish_upload()
                return
            s.px  # completes a pending upload (no-op otherwise)
            launch(s.shard.atom_lo, s.shard.nlocal, self._loc_chunk_off, self._loc_chunk_order, self._loc_nchunk,
                   self._loc_scratch, self.pair_partials)
            return
        _lib.call(
            "hp_promol_weights", self.functor, s.npts, s.px, s.py, s.pz, s.point_base, s.natom,
            s.atom_xyz, s.atom_point_offsets, self.offsets, self.A, self.alpha, self.order,
            self.ntile, self.tiles, s.rho, s.molw, float(density_cutoff), float(promol_offset),
            s.promol if want_promol else None, s.at_w if want_weights else None,
            s.entropy_partials if want_entropy else None, stream_ptr(s.device),
        )  # fmt: skip
