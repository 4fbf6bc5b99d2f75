"""part-dens settings of the product (scripts/partition_density.py: DEFAULTS, CLASS_ARGS) against
the reference's own data/part-dens.yaml and data/keywords.yaml, read where the reference tree is
present.  CPU only."""

import pathlib

import pytest

from horton_part_b200.scripts.partition_density import CLASS_ARGS, DEFAULTS

REF_DATA = pathlib.Path("/root/reference/src/horton_part/data")


@pytest.fixture(scope="module")
def ref_yaml():
    if not REF_DATA.is_dir():
        pytest.skip("reference tree not present on this machine")
    import yaml

    return {name: yaml.safe_load((REF_DATA / name).read_text()) for name in ("part-dens.yaml", "keywords.yaml")}


def test_defaults_equal_part_dens_yaml(ref_yaml):
    assert DEFAULTS == ref_yaml["part-dens.yaml"]


def test_constructor_whitelists_equal_keywords_yaml(ref_yaml):
    keywords = ref_yaml["keywords.yaml"]
    for scheme, mine in CLASS_ARGS.items():
        assert sorted(mine) == sorted(keywords[scheme]["class_args"]), scheme
    # every scheme of the reference's whitelist that partitions on a density is offered
    assert set(CLASS_ARGS) == {k for k, v in keywords.items() if isinstance(v, dict) and "class_args" in v} - {"b", "h", "hi", "mulliken"}
