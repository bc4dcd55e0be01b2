"""CPU: the oracle (both formulations) against the committed golden vectors dumped from the unmodified
reference, and -- when oracle/_ref/ref_overlap is present -- against the reference run live."""
import glob
import os

import numpy as np
import pytest

import datasets
from oracle_lib import LeanOracle, Oracle, edge_checksum, have_reference, run_reference

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


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_lean_oracle_matches_golden(path):
    """oracle/lean_oracle.cpp (the memory-lean three-phase form behind the full-size goldens) against the committed dumps of
    the unmodified reference: read IDs, lengths, frequencies, superReadID, the final edge tuples and the node / edge counts."""
    z = np.load(path)
    lean = LeanOracle(z["bases"], z["offsets"], int(z["min_overlap"]), threads=3).run(keep_edges=True)
    info = lean.read_info()
    assert lean.n_good == int(z["n_good"])
    assert np.array_equal(info["fnv"], z["fnv"]) and np.array_equal(info["len"], z["len"]) and np.array_equal(info["freq"], z["freq"])
    assert np.array_equal(info["sup"], z["sup"])
    assert np.array_equal(lean.edges(), z["edges"])
    c = lean.counters()
    assert c["nodes"] == int(z["number_of_nodes"]) and c["E_final"] == int(z["number_of_edges"]) and c["asymmetric"] == 0
    assert lean.checksum() == edge_checksum(z["edges"])


def test_lean_oracle_matches_port():
    """... and against omega_oracle.cpp (pinned to the reference) on every seeded set of the parity suite, counters included."""
    for cfg in datasets.adversarial() + datasets.small_configs():
        a = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.THREE_PHASE, threads=3)
        b = LeanOracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"], threads=3).run(keep_edges=True)
        ia, ib = a.read_info(), b.read_info()
        for k in ("fnv", "len", "freq", "sup"):
            assert np.array_equal(ia[k], ib[k]), (cfg["name"], k)
        assert np.array_equal(a.edges(), b.edges()), cfg["name"]
        ca, cb = a.counters(), b.counters()
        assert (ca["E_pre"], ca["number_of_edges"], ca["number_of_nodes"], ca["P_e"], ca["T"], ca["active_pivots"], ca["max_degree"], ca["P_c"], ca["C_c"]) == \
               (cb["E_pre"], cb["E_final"], cb["nodes"], cb["P_e"], cb["T"], cb["active_pivots"], cb["max_degree"], cb["P_c"], cb["C_c"]), cfg["name"]
        assert b.checksum() == edge_checksum(a.edges())


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


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_contracted_fixture_is_consistent(name):
    """Fixtures for the next row of SURVEY.md 8(f) (contractCompositePaths + removeDeadEndNodes, OverlapGraph.cpp:211-215),
    dumped from the unmodified reference by tests/golden/make_golden.py: structural invariants of the composite edges
    (OverlapGraph.cpp:702-785 mergeEdges / mergeList) relative to the graph at :210 that the CUDA path reproduces."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    if "c_edges" not in g.files:
        pytest.skip("fixture predates the contracted dump")
    ce, ls, cl = g["c_edges"].astype(np.int64), g["c_list_start"], g["c_lists"].astype(np.int64)
    assert int(g["c_number_of_edges"]) == len(ce)
    assert int(g["c_number_of_nodes"]) == len(set(ce[:, 0].tolist()))
    simple = {(int(s), int(d), int(o), int(t)) for s, d, o, t in g["edges"].tolist()}          # (src, dst, overlapOffset, orientation) at :210
    twin_or = {0: 3, 3: 0, 1: 1, 2: 2}
    have = {}
    for i, (s, d, o, nl, off) in enumerate(ce.tolist()):
        have.setdefault((s, d, o), []).append(i)
    interior = set()
    for i, (s, d, o, nl, off) in enumerate(ce.tolist()):
        lst = cl[ls[i]:ls[i + 1]]
        assert len(lst) == nl
        if nl == 0:
            assert (s, d, off, o) in simple                            # an edge the contraction left alone
        else:
            assert lst[:, 1].sum() < off                               # mergeList: the last hop is overlapOffset - sum
            interior.update(lst[:, 0].tolist())
            # the path s -> r1 -> ... -> rk -> d consists of simple edges of the graph at :210 with these hop offsets
            path = [s] + lst[:, 0].tolist() + [d]
            hops = lst[:, 1].tolist() + [off - int(lst[:, 1].sum())]
            for a, b, h in zip(path[:-1], path[1:], hops):
                assert any((a, b, h, t) in simple for t in range(4)), (name, a, b, h)
        # twin: reverse direction, twin orientation (:841-855), the reads of the list reversed
        cands = have.get((d, s, twin_or[o]), [])
        assert any(cl[ls[j]:ls[j + 1], 0].tolist() == cl[ls[i]:ls[i + 1], 0].tolist()[::-1] for j in cands), (name, s, d)
    # a read inside a composite edge has been contracted away: it owns no edge any more
    assert not (interior & set(ce[:, 0].tolist()))


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_contraction_restatement_matches_reference(name):
    """oracle/contract_oracle.py (plain-Python restatement of OverlapGraph.cpp:211-215), started from the CANONICAL order
    of the graph at :210, reproduces the unmodified reference's composite edges -- endpoints, orientation, offset and the
    read / offset / orientation lists -- on every fixture: the stage does not depend on the reference's list order."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("contract_oracle", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "contract_oracle.py"))
    co = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(co)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    if "c_edges" not in g.files:
        pytest.skip("fixture predates the contracted dump")
    got = co.Graph(g["edges"].tolist(), g["len"].tolist()).simplify().edge_records()
    ce, ls, cl = g["c_edges"].astype(np.int64), g["c_list_start"], g["c_lists"].astype(np.int64)
    want = sorted((int(s), int(d), int(o), int(off), tuple(cl[ls[i]:ls[i + 1], 0].tolist()), tuple(cl[ls[i]:ls[i + 1], 1].tolist()),
                   tuple(cl[ls[i]:ls[i + 1], 2].tolist())) for i, (s, d, o, nl, off) in enumerate(ce.tolist()))
    assert len(got) == len(want), (len(got), len(want))
    assert got == want


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/ref_overlap not built (needs /root/reference)")
def test_contraction_restatement_matches_live_reference(tmp_path):
    """The same comparison against the unmodified reference run here (--dump2), on seeded sets beyond the fixtures:
    every adversarial set plus larger samples of configs 2, 3 and 5."""
    import importlib.util
    from metagenomics_b200 import synth
    spec = importlib.util.spec_from_file_location("contract_oracle", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "contract_oracle.py"))
    co = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(co)
    sets = datasets.adversarial() + [synth.config(2, scale=0.004), synth.config(3, scale=0.0006), synth.containment_stress(9, genome_len=9000, n_primary=2500)]
    for k, cfg in enumerate(sets):
        fa = str(tmp_path / f"in{k}.fa")
        synth.write_fasta(fa, cfg["bases"], cfg["offsets"])
        d, _, _ = run_reference([fa], cfg["min_overlap"], paired=cfg["paired"], contracted=True)
        c = d["contracted"]
        ce, ls, cl = c["edges"].astype(np.int64), c["list_start"], c["lists"].astype(np.int64)
        want = sorted((int(s), int(t), int(o), int(off), tuple(cl[ls[i]:ls[i + 1], 0].tolist()), tuple(cl[ls[i]:ls[i + 1], 1].tolist()),
                       tuple(cl[ls[i]:ls[i + 1], 2].tolist())) for i, (s, t, o, nl, off) in enumerate(ce.tolist()))
        got = co.Graph(d["edges"].tolist(), d["reads"]["len"].tolist()).simplify().edge_records()
        assert got == want, (cfg["name"], len(got), len(want))
