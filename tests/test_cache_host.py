"""Contract of horton_part_b200/core/cache.py, the container behind ``part.cache`` / ``part[key]``
and the ``@just_once`` methods of the WPart API: the behaviours the reference's
tests/core/test_cache.py pins (keys, in-place allocation that survives ``clear``, defaults, tags,
argument errors).  CPU only."""

import numpy as np
import pytest

from horton_part_b200.core.cache import Cache, JustOnceClass, just_once


class Counter(JustOnceClass):
    def __init__(self):
        super().__init__()
        self.n = 0

    @just_once
    def bump(self):
        self.n += 1

    def bump_always(self):
        self.n += 1


def test_just_once_runs_again_only_after_clear():
    c = Counter()
    c.bump()
    c.bump()
    assert c.n == 1
    c.clear()
    assert c.n == 1
    c.bump()
    c.bump()
    assert c.n == 2
    c.bump_always()
    assert c.n == 3


@pytest.mark.parametrize("store", ["dump", "setitem"])
def test_keys_are_names_or_tuples(store):
    c = Cache()
    if store == "dump":
        c.dump("foo", 5)
        c.dump("foo", 4, 6)
        c.dump(("egg",), 7)
    else:
        c["foo"] = 5
        c["foo", 4] = 6
        c[("egg",)] = 7
    for key in ("foo", ("foo",)):
        assert c[key] == 5 and c.load(key) == 5
    assert c["foo", 4] == c[("foo", 4)] == c.load("foo", 4) == c.load(("foo", 4)) == 6
    assert "egg" in c and ("foo", 4) in c and ("foo", 5) not in c
    assert len(c) == 3
    c.clear()
    assert len(c) == 0 and len(c._store) == 0


@pytest.mark.parametrize("shape", [5, (5, 10)])
def test_alloc_returns_the_same_zeroed_array(shape):
    c = Cache()
    assert "bar" not in c
    arr, new = c.load("bar", alloc=shape)
    expected = (shape,) if isinstance(shape, int) else shape
    assert new and (arr == 0).all() and arr.shape == expected and arr.dtype == float
    arr[3] = 1
    assert c.load("bar") is arr
    again, new = c.load("bar", alloc=shape)
    assert not new and again is arr and (again[3] == 1).all()


def test_arrays_survive_clear_and_are_reissued_zeroed():
    c = Cache()
    a1, new = c.load("egg", alloc=(5, 10))
    assert new
    with pytest.raises(TypeError):
        c.load("egg", alloc=10)
    with pytest.raises(TypeError):
        c.load("egg", alloc=(10, 5))
    a1[:] = 1.0
    c.clear()
    assert "egg" not in c and (a1 == 0.0).all()  # invalidated: zeroed in place, not reachable
    with pytest.raises(KeyError):
        c.load("egg")
    a2, new = c.load("egg", alloc=(5, 10))
    assert new and a2 is a1 and "egg" in c and c.load("egg") is a1
    c.clear()
    a3, new = c.load("egg", alloc=(5, 1, 2))  # another shape: a new array
    assert new and a3.shape == (5, 1, 2) and a3 is not a1


def test_defaults():
    c = Cache()
    assert c.load("egg", default=5) == 5
    c.dump("egg", 5)
    assert c.load("egg", default=6) == 5
    c.clear()
    assert c.load("egg", default=6) == 6 and c.load("egg", default=None) is None
    with pytest.raises(KeyError):
        c.load("egg")
    c.dump("floep", np.array([3.1, 5.1]))
    assert (c.load("floep", default=3) == [3.1, 5.1]).all()
    c.clear()
    assert c.load("floep", default=3) == 3


def test_argument_errors():
    c = Cache()
    with pytest.raises(KeyError):
        c.load("boo")
    c.dump("bar", np.zeros(4))
    with pytest.raises(TypeError):
        c.load("bar", alloc=5)  # present, but not as an allocated array of that shape
    for args, kwargs in (((), {}), (("foo",), {"sadfj": 4}), (("foo",), {"alloc": 3, "sdasffd": 0}),
                         (("foo",), {"alloc": 3, "default": 0}), (("foo",), {"jgfjg": 3, "default": 0})):  # fmt: skip
        with pytest.raises(TypeError):
            c.load(*args, **kwargs)
    with pytest.raises(TypeError):
        c.dump()
    with pytest.raises(TypeError):
        c.dump("one")
    with pytest.raises(TypeError):
        c.clear_item()


def test_dealloc_forgets_the_storage():
    c = Cache()
    c.dump("foo", 5)
    c.dump("bar", 6)
    c.clear_item("foo", dealloc=True)
    assert "foo" not in c and "bar" in c and len(c._store) == 1
    c.dump("foo", 5)
    c.clear(dealloc=True)
    assert len(c) == 0 and len(c._store) == 0


def test_iteration_and_tags():
    c = Cache()
    c.dump("foo", 5, tags="c")
    c.dump("bar", 6)
    c.dump("egg", 7, tags="op")
    c.dump("spam", 8, tags="co")
    assert sorted(c) == sorted(c.iterkeys()) == ["bar", "egg", "foo", "spam"] and len(c) == 4
    assert sorted(c.itervalues()) == [5, 6, 7, 8]
    assert sorted(c.iteritems()) == [("bar", 6), ("egg", 7), ("foo", 5), ("spam", 8)]
    assert sorted(c.iterkeys(tags="c")) == ["foo", "spam"] and sorted(c.itervalues(tags="c")) == [5, 8]
    assert sorted(c.iteritems(tags="o")) == [("egg", 7), ("spam", 8)]
    assert list(c.iterkeys(tags="a")) == []
    # what do_all() returns: the keys tagged "o" (core/base.py:404-410)
    assert sorted(c.iterkeys(tags="o")) == ["egg", "spam"]


def test_tags_in_load_and_clear():
    c = Cache()
    c.dump("a", 5, tags="ab")
    for tags in ("a", "ab", "abc"):
        with pytest.raises(TypeError):
            c.load("a", tags=tags)  # tags only go with alloc
    with pytest.raises(TypeError):
        c.load("b", default=5, tags="abc")
    c.clear(tags="cd")
    assert len(c) == 1
    c.clear(tags="a")
    assert len(c) == 0
    t1, new = c.load("tmp", alloc=5, tags="qw")
    t2, again = c.load("tmp", alloc=5, tags="qw")
    assert new and not again and t1 is t2
    for tags in ("w", "aw"):
        with pytest.raises(ValueError):
            c.load("tmp", alloc=5, tags=tags)


def test_deferred_values_are_assembled_on_first_access():
    from horton_part_b200.core.cache import Cache, Deferred

    calls = []

    def build():
        calls.append(1)
        return np.arange(4.0)

    c = Cache()
    c.dump("whole", Deferred(build), tags="o")
    assert "whole" in c and not calls  # membership does not assemble
    assert np.array_equal(c.load("whole"), np.arange(4.0)) and len(calls) == 1
    assert c.load("whole") is c.load("whole") and len(calls) == 1  # assembled once, then an ordinary value
    c.dump("never", Deferred(build))
    c.clear()
    assert len(calls) == 1  # clearing an unassembled value does not build it
    with __import__("pytest").raises(KeyError):
        c.load("never")
    c.dump("again", Deferred(build))
    assert dict(c.iteritems())["again"].shape == (4,) and len(calls) == 2


def test_dump_many_equals_repeated_dump():
    from horton_part_b200.core.cache import Cache

    a, b = Cache(), Cache()
    items = [(f"k_{i}", np.full(3, float(i))) for i in range(5)]
    for k, v in items:
        a.dump(k, v, tags="o")
    b.dump_many(iter(items), tags="o")
    assert sorted(a.iterkeys(tags="o")) == sorted(b.iterkeys(tags="o")) == sorted(k for k, _ in items)
    for k, v in items:
        assert b.load(k) is v
        assert b.load(k, alloc=3, tags="o")[1] is False
    b.clear(tags="o")
    assert "k_0" not in b
