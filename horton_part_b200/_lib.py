"""ctypes binding of the C-ABI library ``libhp_b200.so`` (declared in include/hp_b200.h).

There is no CPU fallback: if the library is missing, cannot be loaded, or a call fails, an
exception is raised.  ``call(name, *args)`` converts torch tensors to raw device pointers and maps
non-zero status codes to ``RuntimeError`` / ``ValueError`` carrying ``hp_last_error()``.
"""

from __future__ import annotations

import ctypes as C
import pathlib

__all__ = ["lib", "call", "ptr", "EXPORTS", "library_path", "HpError"]

_PKG = pathlib.Path(__file__).resolve().parent
import os as _os

_LIBPATH = pathlib.Path(_os.environ.get("HP_B200_LIB", _PKG / "libhp_b200.so"))  # override: tuning builds only

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double
_int = C.c_int
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol include/hp_b200.h declares
EXPORTS = {
    "hp_last_error": (C.c_char_p, []),
    "hp_abi_version": (_int, []),
    "hp_device_props": (_int, [_p]),
    "hp_num_partials": (_i32, []),
    "hp_local_index_scratch_bytes": (_sz, [_i64]),
    "hp_build_local_index": (_int, [_p, _i64, _p, _f64, _i64, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "hp_split_points": (_int, [_p, _i64, _p, _p, _p, _p]),
    "hp_table_mbis": (_int, [_i32, _p, _p, _p, _p]),
    "hp_table_scaled": (_int, [_i32, _p, _p, _p, _p]),
    "hp_table_nlis": (_int, [_i32, _p, _p, _p, _p, _p, _p]),
    "hp_tile_limits": (None, [_p, _p]),
    "hp_promol_weights": (
        _int,
        [_int, _i64, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _f64, _f64, _p, _p, _p, _p],
    ),
    "hp_promol_weights_local": (
        _int,
        [_int, _i64, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _f64, _f64, _f64, _p, _f64,
         _i32, _i32, _p, _p, _i64, _p, _p, _p, _p, _p, _p],
    ),
    "hp_local_tile_limits": (None, [_p, _p]),
    "hp_local_chunk_points": (_i32, []),
    "hp_shell_screen": (_int, [_i32, _i32, _p, _p, _p, _f64, _p, _p]),
    "hp_shell_project": (_int, [_i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hp_shell_harmonics": (_int, [_i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hp_mbis_radial_solve": (
        _int,
        [_i32, _i32, _p, _p, _p, _p, _p, _p, _p, _f64, _f64, _i32, _i32, _i32, _p, _p, _p, _p, _p],
    ),
    "hp_nlis_radial_solve": (
        _int,
        [_i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f64, _f64, _i32, _i32, _i32, _p, _p, _p, _p, _p],
    ),
    "hp_lisa_sc_radial_solve": (
        _int,
        [_i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _f64, _f64, _f64, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p],
    ),
    "hp_spline_build": (_int, [_i32, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _p]),
    "hp_spline_system_inverse": (_int, [_i32, _p, _p]),
    "hp_spline_lut_size": (_i32, [_i32, _p]),
    "hp_spline_lut_fill": (_int, [_i32, _p, _p, _p]),
    "hp_spline_tile_limits": (None, [_p, _p]),
    "hp_promol_weights_spline": (
        _int,
        [_i64, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _f64, _f64, _p, _p, _f64, _p, _p, _p, _p],
    ),
    "hp_isa_update": (_int, [_i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hp_spline_integral_blocks": (_i32, [_i64]),
    "hp_atom_weight_integrals_spline": (
        _int,
        [_i64, _p, _p, _p, _i32, _p, _p, _p, _p, _f64, _p, _p, _p, _p, _p, _p],
    ),
    "hp_molgrid_num_blocks": (_i32, [_i64]),
    "hp_shell_moments": (
        _int,
        [_int, _i64, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _f64, _i32, _i32, _p, _p, _p],
    ),
    "hp_atom_weight_integrals": (
        _int,
        [_int, _i64, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p],
    ),
    "hp_radial_change": (_int, [_i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hp_radial_valid": (_int, [_i32, _i32, _p, _p, _p, _p, _p, _i32, _i32, _f64, _i32, _p, _p]),
    "hp_molgrid_update_tile_limits": (None, [_p, _p]),
    "hp_molgrid_update_pass": (
        _int,
        [_int, _i64, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _f64, _i32, _p, _p, _p],
    ),
    "hp_hessian_scratch_bytes": (_sz, [_i32]),
    "hp_hessian": (
        _int,
        [_int, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f64, _i32, _p, _sz, _p, _p],
    ),
    "hp_finish_iteration": (_int, [_i32, _p, _i32, _p, _p, _p]),
    "hp_sum_partials": (_int, [_i32, _p, _p, _p]),
    "hp_segment_integrate": (_int, [_i32, _p, _p, _p, _p, _p, _p]),
    "hp_atom_moments": (_int, [_i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hp_becke_weights": (_int, [_i64, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _p]),
    "hp_neighbor_counts": (_int, [_i32, _p, _p, _i32, _i32, _p, _p, _p]),
    "hp_host_is_pinned": (_int, [_p]),
    "hp_host_to_device": (_int, [_p, _p, _sz, _p, _sz, _i32, _p]),
    "hp_dfma_probe": (_int, [_i32, _p, _p, _p, _p]),
    "hp_shell_moments_table": (
        _int,
        [_i64, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f64, _i32, _i32, _i32, _p, _p, _p],
    ),
    "hp_hessian_table": (
        _int,
        [_i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f64, _i32, _p, _sz, _p, _p],
    ),
    "hp_fold_chunk_entropy": (_int, [_i64, _p, _p, _p]),
    "hp_hessian_tiles_executed": (_int, [_i32, _i64, _p, _p, _p, _p, _p]),
    "hp_aim_on_points": (_int, [_int, _i64, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _f64, _p, _p, _p, _p]),
    "hp_comm_nccl_version": (_i32, []),
    "hp_comm_unique_id": (_int, [_p]),
    "hp_comm_init": (_int, [_i32, _i32, _p, _p]),
    "hp_comm_allreduce": (_int, [_p, _p, _i64, _i32, _p]),
    "hp_comm_allgather": (_int, [_p, _p, _p, _i64, _p]),
    "hp_comm_destroy": (_int, [_p]),
    "hp_loop_begin": (_int, [_p, _p]),
    "hp_loop_stamp": (_int, [_p, _p, _i32, _i32, _p, _p]),
    "hp_loop_commit": (_int, [_p, _i32, _p, _p, _f64, _i32, _p, _p, _p, _p]),
    "hp_loop_end": (_int, [_p, _p]),
    "hp_loop_launch": (_int, [_p, _p]),
    "hp_loop_destroy": (_int, [_p, _p]),
}


# int-returning functions whose result is a value, not a status code
_NOT_STATUS = {"hp_abi_version", "hp_num_partials", "hp_local_index_scratch_bytes", "hp_last_error", "hp_tile_limits", "hp_molgrid_num_blocks", "hp_hessian_scratch_bytes",
               "hp_molgrid_update_tile_limits", "hp_host_is_pinned", "hp_local_tile_limits", "hp_local_chunk_points",
               "hp_spline_integral_blocks", "hp_spline_lut_size", "hp_spline_tile_limits", "hp_comm_nccl_version"}


class HpError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


_lib = None


def library_path() -> pathlib.Path:
    return _LIBPATH


def lib():
    """Load (once) and return the ctypes handle; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not _LIBPATH.exists():
            raise ImportError(
                f"{_LIBPATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)"
            )
        handle = C.CDLL(str(_LIBPATH))
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def ptr(t):
    """Raw pointer of a torch tensor / None / int (streams)."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    if hasattr(t, "ctypes"):  # host numpy array for *_host parameters
        return t.ctypes.data
    raise TypeError(f"cannot pass {type(t)} to the C ABI")


def call(name: str, *args):
    handle = lib()
    fn = getattr(handle, name)
    conv = [ptr(a) if (at is _p) else a for a, at in zip(args, fn.argtypes)]
    if len(conv) != len(fn.argtypes):
        raise TypeError(f"{name} expects {len(fn.argtypes)} arguments, got {len(args)}")
    rc = fn(*conv)
    if name not in _NOT_STATUS and rc != 0:
        msg = handle.hp_last_error().decode(errors="replace")
        if rc == 1:
            raise ValueError(f"{name}: {msg}")
        raise HpError(f"{name} failed (status {rc}): {msg}")
    return rc
