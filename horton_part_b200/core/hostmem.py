"""Host-side buffers of the device path: page-locked staging, uploads from NumPy arrays, downloads
into pooled page-locked arrays.

The class API takes and returns NumPy arrays (pageable memory).  At config 5 the slab is 2.8 GB up
and 0.93 GB down per job, so how those bytes move decides the end-to-end rate:

* ``upload``      NumPy -> device.  Page-locked sources (``pinned_empty``) go out as one async copy;
                  pageable sources are pipelined through a pooled page-locked staging buffer by
                  ``hp_host_to_device`` (multi-threaded memcpy overlapped with the DMA).
* ``download``    device -> a pooled page-locked buffer, returned as a NumPy array without a second
                  copy.  The pool re-issues a buffer only when nothing refers to it any more; idle
                  buffers of other sizes are kept up to HP_B200_PINNED_POOL_BYTES (default 8 GiB).
* ``pinned_empty``  for callers who want their input arrays page-locked from the start.
"""

from __future__ import annotations

import os

import numpy as np

from .. import _lib

__all__ = ["pinned_empty", "is_pinned", "upload", "upload_into", "download", "pool_stats"]

_STAGING_BYTES = 128 << 20
_staging = None
_pending = []  # (device, tensor, array) of asynchronous uploads from page-locked sources not yet drained
_pool = []  # page-locked uint8 tensors handed out by download()
_stats = {"pinned_allocs": 0, "pinned_reuses": 0, "staged_uploads": 0, "direct_uploads": 0}


def _threads():
    return max(1, min(16, (os.cpu_count() or 2) // 2))


def pinned_empty(shape, dtype=np.float64):
    """Uninitialised page-locked NumPy array (owned by a torch tensor kept alive through
    ``ndarray.base``): uploads from it run at PCIe speed."""
    import torch

    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    buf = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True)
    return buf.numpy()[:n].view(dtype).reshape(shape)


def is_pinned(array) -> bool:
    return bool(_lib.call("hp_host_is_pinned", int(array.ctypes.data)))


def _staging_buffer():
    global _staging
    if _staging is None:
        import torch

        _staging = torch.empty(_STAGING_BYTES, dtype=torch.uint8, pin_memory=True)
    return _staging


def upload_into(out, array, stream=None):
    """Copy the C-contiguous NumPy ``array`` into the device tensor ``out`` (same number of bytes) on the CUDA
    stream ``stream`` (raw handle; default: the device's current stream).  Page-locked sources go out as one
    asynchronous copy and stay referenced until :func:`drain`; pageable ones are pipelined through the pooled
    page-locked staging buffer (the call returns when the source has been read)."""
    import torch

    a = np.ascontiguousarray(array)
    nbytes = a.nbytes
    if nbytes != out.numel() * out.element_size() or not out.is_contiguous():
        raise ValueError("upload_into: destination must be a contiguous tensor of the source's size")
    if nbytes == 0:
        return out
    device = out.device
    if stream is None:
        stream = torch.cuda.current_stream(device).cuda_stream
    if is_pinned(a):
        # asynchronous copy straight from the caller's page-locked array: the source stays referenced in
        # _pending until drain() has synchronised the stream the consumers run on (GridSlab does that once for
        # all of its uploads), so neither the garbage collector nor torch's host allocator can recycle it while
        # the DMA engine still reads it.  The caller must not modify the array before drain() returns.
        _stats["direct_uploads"] += 1
        _lib.call("hp_host_to_device", out, int(a.ctypes.data), nbytes, None, 0, 1, stream)
        _pending.append((torch.device(device), out, a))
        return out
    stage = _staging_buffer()
    _stats["staged_uploads"] += 1
    _lib.call("hp_host_to_device", out, int(a.ctypes.data), nbytes, stage, stage.numel(), _threads(), stream)
    return out


def upload(array, device, dtype=None):
    """Device tensor holding a copy of ``array`` (C-contiguous, optionally converted to ``dtype``)."""
    import torch

    a = np.ascontiguousarray(array, dtype=dtype)
    src = torch.from_numpy(a)
    if a.nbytes < (8 << 20):
        return src.to(device)
    return upload_into(torch.empty(src.shape, dtype=src.dtype, device=device), a)


def drain(device=None):
    """Wait for the asynchronous uploads issued by :func:`upload` (on ``device``'s current stream) and
    release the references that kept their page-locked sources alive."""
    import torch

    if not _pending:
        return
    devices = {d for d, _, _ in _pending} if device is None else {torch.device(device)}
    for d in devices:
        torch.cuda.current_stream(d).synchronize()
    _pending[:] = [p for p in _pending if p[0] not in devices]


def _take(nbytes):
    """Pool entry [buffer, weakref-to-the-array-handed-out] whose array is gone (or a new one)."""
    import torch

    for entry in _pool:
        buf, ref = entry
        if ref is not None and ref() is not None:
            continue  # the array (or a view of it: views keep their base alive) is still in use
        if nbytes <= buf.numel() <= 2 * nbytes + 4096:
            _stats["pinned_reuses"] += 1
            return entry
    # No idle buffer fits.  Idle buffers of other sizes stay (freeing and re-allocating page-locked memory costs
    # ~0.1 s per GB: a process that alternates between job sizes would pay it on every call); the oldest idle
    # ones go only when the pool would exceed its budget.
    budget = int(os.environ.get("HP_B200_PINNED_POOL_BYTES", 8 << 30))
    total = sum(e[0].numel() for e in _pool) + nbytes
    if total > budget:
        keep = []
        for e in _pool:  # oldest first
            idle = e[1] is None or e[1]() is None
            if idle and total > budget:
                total -= e[0].numel()
            else:
                keep.append(e)
        _pool[:] = keep
    entry = [torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True), None]
    _stats["pinned_allocs"] += 1
    _pool.append(entry)
    return entry


def download(tensor):
    """NumPy array with the contents of a device tensor.  Large tensors land in a pooled page-locked
    buffer (no second copy).  The buffer is re-issued only after the returned array and every view
    of it have been garbage-collected."""
    import weakref

    import torch

    t = tensor.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (8 << 20):
        return t.cpu().numpy()
    entry = _take(nbytes)
    host = entry[0][:nbytes].view(t.dtype).reshape(t.shape)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    array = host.numpy()
    entry[1] = weakref.ref(array)
    return array


def pool_stats():
    return dict(_stats, pool_buffers=len(_pool), pool_bytes=sum(e[0].numel() for e in _pool))
