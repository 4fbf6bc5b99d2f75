"""Result store and run-once bookkeeping of the partitioning classes (host side).

Mirrors the behaviour of the reference's ``Cache`` / ``JustOnceClass`` / ``just_once``
(/root/reference/src/horton_part/core/cache.py:34-80, 190-399): keyed items with tag sets and a
validity flag, ``load(key, alloc=shape)`` returning ``(array, is_new)``, ``dump``, tag-filtered
iteration, ``clear`` that zeroes arrays in place and marks them stale so that the next
``alloc`` load reuses the memory.  This is pure host bookkeeping: device buffers never live here,
only the NumPy results downloaded from them.
"""

import time

import numpy as np

__all__ = ["JustOnceClass", "just_once", "Cache"]


class JustOnceClass:
    """Base class remembering which ``@just_once`` methods already ran (until ``clear()``)."""

    def __init__(self):
        self._done_just_once = set()
        self.time_usage = {}

    def __clear__(self):
        self.clear()

    def clear(self):
        self._done_just_once = set()


def just_once(method):
    """Run ``method`` only on the first call per instance; record its wall time in ``time_usage``."""

    def run_once(self):
        done = getattr(self, "_done_just_once", None)
        if done is None:
            raise TypeError(
                "Missing hidden _done_just_once. Forgot to call JustOnceClass.__init__()?"
            )
        name = method.__name__
        if name in done:
            return None
        start = time.time()
        method(self)
        self.time_usage[name] = time.time() - start
        done.add(name)
        return None

    run_once.__doc__ = method.__doc__
    run_once.__name__ = method.__name__
    return run_once


def _shape_of(alloc):
    shape = alloc if hasattr(alloc, "__len__") else (alloc,)
    if len(shape) == 0:
        raise TypeError("Alloc can not be an empty list")
    return tuple(shape)


def _tagset(tags):
    return set() if tags is None else set(tags)


class Deferred:
    """A cache value that is assembled on first access: ``Deferred(thunk)`` stored with :meth:`Cache.dump`
    is replaced by ``thunk()`` the first time the key is loaded or iterated over.  Used for results whose
    data is already on the host but whose reference-shaped array costs time to build (a sharded rank's
    whole-grid ``promoldens``: 0.5 GB of pages for a slice that is already there)."""

    __slots__ = ("thunk",)

    def __init__(self, thunk):
        self.thunk = thunk


class _Entry:
    __slots__ = ("_value", "valid", "tags")

    def __init__(self, value, tags):
        self._value = value
        self.valid = True
        self.tags = _tagset(tags)

    @property
    def value(self):
        if isinstance(self._value, Deferred):
            self._value = self._value.thunk()
        return self._value

    @value.setter
    def value(self, v):
        self._value = v

    def matches(self, shape):
        v = self.value
        return isinstance(v, np.ndarray) and v.shape == shape and issubclass(v.dtype.type, float)

    def invalidate(self):
        """Returns True when the payload could be wiped in place (memory kept for reuse)."""
        self.valid = False
        if isinstance(self._value, Deferred):
            return False  # never assembled: nothing to keep
        v = self.value
        if isinstance(v, np.ndarray):
            v[...] = 0.0
            return True
        wipe = getattr(v, "__clear__", None)
        if callable(wipe):
            wipe()
            return True
        return False


_MISSING = object()


def _flatten_key(key):
    if hasattr(key, "__len__") and len(key) == 0:
        raise TypeError("At least one argument needed to specify a key.")
    while isinstance(key, tuple) and len(key) == 1:
        key = key[0]
    return key


class Cache:
    """Dictionary-like store of computed results with tags, validity and in-place reuse."""

    def __init__(self):
        self._store = {}

    # -- writing -------------------------------------------------------------------------------
    def dump(self, *args, **kwargs):
        tags = kwargs.pop("tags", None)
        if kwargs:
            raise TypeError("Unknown optional arguments: %s" % list(kwargs))
        if len(args) < 2:
            raise TypeError("At least two arguments are required: key1 and value.")
        self._store[_flatten_key(args[:-1])] = _Entry(args[-1], tags)

    def dump_many(self, items, tags=None):
        """``dump(key, value, tags=tags)`` for every (key, value) of ``items`` (keys already flat): the per-atom
        results of a large system are thousands of entries, and the per-call bookkeeping of ``dump`` adds up."""
        tagset = _tagset(tags)  # shared by the entries: tag sets are only compared and intersected
        store = self._store
        for key, value in items:
            entry = _Entry.__new__(_Entry)
            entry._value, entry.valid, entry.tags = value, True, tagset
            store[key] = entry

    def load(self, *key, **kwargs):
        key = _flatten_key(key)
        alloc = kwargs.pop("alloc", None)
        default = kwargs.pop("default", _MISSING)
        tags = kwargs.pop("tags", None)
        if alloc is not None and default is not _MISSING:
            raise TypeError(
                "The optional arguments alloc and default can not be used at the same time."
            )
        if tags is not None and alloc is None:
            raise TypeError("The tags argument is only allowed when the alloc argument is present.")
        if kwargs:
            raise TypeError("Unknown optional arguments: %s" % list(kwargs))

        entry = self._store.get(key)
        if alloc is None:
            if entry is not None and entry.valid:
                return entry.value
            if default is not _MISSING:
                return default
            raise KeyError(key)

        shape = _shape_of(alloc)
        if entry is None:
            entry = self._store[key] = _Entry(np.zeros(shape, dtype=float), tags)
            return entry.value, True
        if entry.valid:
            if not entry.matches(shape):
                raise TypeError("The stored item does not match the given alloc.")
            if _tagset(tags) != entry.tags:
                raise ValueError("Tags do not match.")
            return entry.value, False
        # stale entry: reuse its memory when it has the right shape, else replace it
        if entry.matches(shape):
            entry.valid = True
            if _tagset(tags) != entry.tags:
                raise ValueError("Tags do not match.")
        else:
            entry = self._store[key] = _Entry(np.zeros(shape, dtype=float), tags)
        return entry.value, True

    # -- clearing ------------------------------------------------------------------------------
    def clear(self, **kwargs):
        dealloc = kwargs.pop("dealloc", False)
        tags = _tagset(kwargs.pop("tags", None))
        if kwargs:
            raise TypeError("Unexpected arguments: %s" % list(kwargs))
        for key, entry in list(self._store.items()):
            if not tags or (entry.tags & tags):
                self.clear_item(key, dealloc=dealloc)

    def clear_item(self, *key, **kwargs):
        key = _flatten_key(key)
        dealloc = kwargs.pop("dealloc", False)
        if kwargs:
            raise TypeError("Unexpected arguments: %s" % list(kwargs))
        entry = self._store.get(key)
        if entry is None:
            return
        if dealloc or not entry.invalidate():
            del self._store[key]

    # -- reading -------------------------------------------------------------------------------
    def __contains__(self, key):
        entry = self._store.get(_flatten_key(key))
        return entry is not None and entry.valid

    def __len__(self):
        return sum(1 for e in self._store.values() if e.valid)

    def __getitem__(self, key):
        return self.load(key)

    def __setitem__(self, key, value):
        self.dump(key, value)

    def __iter__(self):
        return self.iterkeys()

    def _select(self, tags):
        tags = _tagset(tags)
        for key, entry in self._store.items():
            if entry.valid and (not tags or (entry.tags & tags)):
                yield key, entry.value

    def iterkeys(self, tags=None):
        return (k for k, _ in self._select(tags))

    def itervalues(self, tags=None):
        return (v for _, v in self._select(tags))

    def iteritems(self, tags=None):
        return self._select(tags)
