"""CPU test: the product's direct table construction (ic3net_b200/traffic_helper.py)
against the tables the unmodified reference builds (tests/golden/tj_tables_*.npz),
plus the reference's own route self-check (traffic_junction_env.py:526-537)."""
import numpy as np
import pytest

from helpers import golden_names, load_golden
from ic3net_b200 import traffic_helper as th


@pytest.mark.parametrize("name", golden_names("tj_tables_"))
def test_tables_match_reference(name):
    meta, z = load_golden(name)
    t = th.build_tables(meta["difficulty"], meta["dim"])
    assert list(t["dims"]) == meta["dims"]
    assert (t["BASE"], t["OUTSIDE"], t["CAR"], t["vocab"], t["npath"]) == \
        (meta["BASE"], meta["OUTSIDE"], meta["CAR"], meta["vocab"], meta["npath"])
    assert np.array_equal(t["grid"], z["grid"])
    ln, cells = z["route_len"], z["route_cells"]
    assert np.array_equal(t["route_len"], ln)
    for g in range(ln.shape[0]):
        for p in range(ln.shape[1]):
            assert np.array_equal(t["routes"][g][p], cells[g, p, :ln[g, p]])
            packed = t["route_cells"][g, p, :ln[g, p]]
            assert np.array_equal(packed >> 16, cells[g, p, :ln[g, p], 0])
            assert np.array_equal(packed & 0xffff, cells[g, p, :ln[g, p], 1])


def test_dim_asserts():
    with pytest.raises(AssertionError):
        th.build_tables("medium", 7)
    with pytest.raises(AssertionError):
        th.build_tables("hard", 10)
    with pytest.raises(AssertionError):
        th.build_tables("hard", 6)


def test_easy_quirk_ids_alias_outside():
    # traffic_junction_env.py:112-124: dims grow to D+1 but BASE uses D -> ids 12, 13 exist with OUTSIDE = 12
    t = th.build_tables("easy", 6)
    assert t["OUTSIDE"] == 12 and t["grid"].max() == 13 and t["vocab"] == 15
