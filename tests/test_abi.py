"""The C-ABI library loads and exports every symbol include/hp_b200.h declares (no compute)."""

import re

from conftest import ROOT


def test_header_symbols_exported(built_lib):
    from horton_part_b200 import _lib

    header = (ROOT / "include" / "hp_b200.h").read_text()
    declared = set(re.findall(r"HP_API[^;(]*?\b(hp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name)
    assert _lib.call("hp_abi_version") == 1
    assert _lib.call("hp_num_partials") >= 148


def test_value_returning_symbols_are_not_treated_as_status_codes():
    """Entry points whose C return type is not plain ``int`` return a value (a count, a size, a
    string); _lib.call must not map their non-zero results to exceptions."""
    from horton_part_b200 import _lib

    header = (ROOT / "include" / "hp_b200.h").read_text()
    valued = set(re.findall(r"HP_API\s+(?:int32_t|size_t|const char\*|void)\s+(hp_[a-z0-9_]+)\s*\(", header))
    assert valued and valued <= _lib._NOT_STATUS, valued - _lib._NOT_STATUS


def test_argument_errors_map_to_exceptions(built_lib):
    import pytest

    from horton_part_b200 import _lib

    with pytest.raises(ValueError, match="hp_table_mbis"):
        _lib.call("hp_table_mbis", 0, None, None, None, None)
    assert b"bad arguments" in _lib.lib().hp_last_error()


def test_no_cpu_fallback():
    """Without CUDA the product refuses to run instead of falling back."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from horton_part_b200.core.device import require_cuda

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        require_cuda()
