"""CPU: the C-ABI library loads, exports every symbol include/ogb.h declares, and its device entry
points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ogb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ogb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from metagenomics_b200._lib import LIB_PATH, PROTOTYPES, lib
    names = declared_symbols()
    assert len(names) >= 35
    raw = C.CDLL(LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/ogb.h but not exported by libogb.so"
    assert set(names) == set(PROTOTYPES), set(names) ^ set(PROTOTYPES)
    assert lib().ogb_version() == 210


def test_struct_layout():
    from metagenomics_b200._lib import Edge, Stats
    assert C.sizeof(Edge) == 12 and Edge.offset.offset == 8 and Edge.orient.offset == 10
    from metagenomics_b200._lib import SimplifyStats
    from metagenomics_b200.api import CEDGE_DTYPE, CITEM_DTYPE
    assert C.sizeof(SimplifyStats) == 5 * 8 + 4 * 4 + 5 * 4 + 4 and CEDGE_DTYPE.itemsize == 40 and CITEM_DTYPE.itemsize == 8
    assert C.sizeof(Stats) == 17 * 8 + 4 * 4 + 11 * 4 + 20 * 4 + 20 * 4 + 4      # 11 + 20 floats, 20 counters, padded to the 8-byte alignment of the struct


def test_device_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from metagenomics_b200 import Context, OgbError
    with pytest.raises(OgbError) as e:
        Context(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under metagenomics_b200/ or include/ may name it."""
    bad = []
    for base in ("metagenomics_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"liboracle|omega_oracle|oracle_lib|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
