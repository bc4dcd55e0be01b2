"""CPU: the oracle (both formulations) against the committed golden vectors dumped from the unmodified
reference, and -- when oracle/_ref/ref_overlap is present -- against the reference run live."""
import glob
import os

import numpy as np
import pytest

import datasets
from oracle_lib import Oracle, have_reference, run_reference

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def test_fixtures_present():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
@pytest.mark.parametrize("mode", [Oracle.BFS, Oracle.THREE_PHASE], ids=["bfs", "three_phase"])
def test_oracle_matches_golden(path, mode):
    z = np.load(path)
    orc = Oracle(z["bases"], z["offsets"], int(z["min_overlap"])).run_all(mode, threads=2)
    info = orc.read_info()
    assert orc.n_good == int(z["n_good"])                      # Dataset::numberOfReads after the filter
    assert np.array_equal(info["fnv"], z["fnv"])                # sort + dedupe order -> read IDs
    assert np.array_equal(info["len"], z["len"]) and np.array_equal(info["freq"], z["freq"])
    assert np.array_equal(info["sup"], z["sup"])                # markContainedReads
    assert np.array_equal(orc.edges(), z["edges"])              # graph at OverlapGraph.cpp:210
    c = orc.counters()
    assert c["number_of_nodes"] == int(z["number_of_nodes"]) and c["number_of_edges"] == int(z["number_of_edges"])


def test_formulations_agree_on_pre_reduction_edges():
    for cfg in datasets.adversarial() + datasets.small_configs()[:2]:
        a = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
        b = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.THREE_PHASE, threads=3)
        assert np.array_equal(a.edges(pre=True), b.edges(pre=True)), cfg["name"]
        assert np.array_equal(a.edges(), b.edges()), cfg["name"]


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/ref_overlap not built (needs /root/reference)")
def test_oracle_matches_live_reference(tmp_path):
    from metagenomics_b200 import synth
    for cfg in [datasets.even_h(), datasets.one_window(), datasets.tandem(), synth.config(3, scale=0.0008)]:
        fa = str(tmp_path / "in.fa")
        synth.write_fasta(fa, cfg["bases"], cfg["offsets"])
        d, t, table = run_reference([fa], cfg["min_overlap"], paired=cfg["paired"], want_table=True)
        orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
        info = orc.read_info()
        assert np.array_equal(info["fnv"], d["reads"]["fnv"]) and np.array_equal(info["sup"], d["reads"]["sup"])
        assert np.array_equal(orc.edges(), d["edges"]), cfg["name"]
        h = orc.l.oracle_hash_string_length(orc.h)
        for i in range(1, orc.n + 1, 17):                       # HashTable::getListOfReads content
            f, r = orc.get_read(i), orc.get_read(i, True)
            for k, key in enumerate([f[:h], f[-h:], r[:h], r[-h:]]):
                assert np.array_equal(orc.lookup(key), table[(i - 1) * 4 + k])
