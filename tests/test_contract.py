"""Simplification stage (OverlapGraph.cpp:211-215: contractCompositePaths + removeDeadEndNodes to the fix-point), CPU side:

* oracle/contract_rounds.py -- the parallel "priority rounds" formulation in plain Python -- against the sequential restatement
  (oracle/contract_oracle.py) and the unmodified reference's dump on every fixture and on seeded sets;
* tests/contract_emul.cpp -- the per-thread bodies of metagenomics_b200/csrc/ogb_contract.cuh (the functions the CUDA kernels
  call), run thread by thread on the CPU, forwards and backwards -- against the same fixtures.

The kernels themselves are checked by the -m gpu tests (tests/test_gpu_parity.py)."""
import glob
import os

import numpy as np
import pytest

import datasets
from oracle_lib import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))

from contract_lib import build_emul, check_twins, composite_records, fixture_records, load_oracle_module as _load


@pytest.fixture(scope="module")
def emul():
    return build_emul()


@pytest.mark.parametrize("name", FIXTURES)
def test_rounds_formulation_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    if "c_edges" not in g.files:
        pytest.skip("fixture predates the contracted dump")
    cr = _load("contract_rounds")
    assert cr.Graph(g["edges"].tolist(), g["len"].tolist()).simplify().edge_records() == fixture_records(g)


@pytest.mark.parametrize("name", FIXTURES)
def test_kernel_bodies_match_reference(name, emul):
    """The device functions of csrc/ogb_contract.cuh, executed on the CPU in both thread orders, reproduce the unmodified
    reference's graph after its fix-point: end points, orientation, offset, and the three lists of every composite edge."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    if "c_edges" not in g.files:
        pytest.skip("fixture predates the contracted dump")
    want = fixture_records(g)
    for reverse in (False, True):
        edges, items, stats = emul(g["edges"], g["len"], reverse)
        assert composite_records(edges, items) == want, (name, reverse)
        check_twins(edges)
        assert (np.diff(edges["src"].astype(np.int64)) >= 0).all()


def seeded_sets():
    from metagenomics_b200 import synth
    return datasets.adversarial() + [synth.config(2, scale=0.01), synth.config(3, scale=0.002), synth.config(1, scale=1.0),
                                     synth.containment_stress(9, genome_len=9000, n_primary=2500), synth.config(4, scale=0.0006), datasets.paired_mixed()]


def test_seeded_sets_all_formulations_agree(emul):
    """Beyond the fixtures (parallel chains between the same end nodes, tandem repeats, palindromes, the config samples): sequential
    restatement == rounds formulation == kernel bodies."""
    co, cr = _load("contract_oracle"), _load("contract_rounds")
    for cfg in seeded_sets():
        o = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.THREE_PHASE, threads=8)
        e, L = o.edges(), o.read_info()["len"]
        want = co.Graph(e.tolist(), L.tolist()).simplify().edge_records()
        assert cr.Graph(e.tolist(), L.tolist()).simplify().edge_records() == want, cfg["name"]
        edges, items, stats = emul(e, L)
        assert composite_records(edges, items) == want, cfg["name"]
        check_twins(edges)


def test_empty_graph(emul):
    edges, items, stats = emul(np.zeros((0, 4), dtype=np.int64), np.array([100, 100, 100]))
    assert len(edges) == 0 and len(items) == 0


def random_graph(rng, n, m, p_chain):
    """A random overlap-like multigraph with consistent twins (all reads 100 bp, so twin offset == offset): chains of degree-2
    nodes, hubs, triangles, parallel paths, self-loops of both kinds."""
    TW = [3, 1, 2, 0]
    edges, seen = [], set()

    def add(s, d, off, t):
        if s == d and t in (1, 2):
            key = (s, d, off, t)
            if key in seen:
                return
            seen.add(key)
            edges.extend([(s, d, off, t), (s, d, off, t)])
            return
        a, b = (s, d, off, t), (d, s, off, TW[t])
        if a in seen or b in seen:
            return
        seen.update((a, b))
        edges.extend([a, b])
    perm = rng.permutation(n) + 1
    k = 0
    while k + 1 < n and rng.random() < p_chain:                                    # paths with a consistent direction: contractible chains
        ln = int(rng.integers(2, 12))
        t = int(rng.choice([0, 3]))
        for j in range(k, min(k + ln, n - 1)):
            add(int(perm[j]), int(perm[j + 1]), int(rng.integers(1, 50)), t)
        k += ln + int(rng.integers(0, 2))
    for _ in range(m):
        s, d = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
        if s == d and rng.random() < 0.7:
            continue
        add(s, d, int(rng.integers(1, 50)), int(rng.integers(0, 4)))
    return np.array(edges, dtype=np.int64).reshape(-1, 4)[:, [0, 1, 2, 3]], np.full(n, 100)


def test_random_graphs_all_formulations_agree(emul):
    co, cr = _load("contract_oracle"), _load("contract_rounds")
    rng = np.random.default_rng(20261018)
    merged = 0
    for it in range(400):
        n = int(rng.integers(3, 60))
        e, L = random_graph(rng, n, int(rng.integers(0, n)), float(rng.uniform(0.3, 0.98)))
        if len(e) == 0:
            continue
        tup = [(int(s), int(d), int(o), int(t)) for s, d, o, t in e.tolist()]
        want = co.Graph(tup, L.tolist()).simplify().edge_records()
        assert cr.Graph(tup, L.tolist()).simplify().edge_records() == want, it
        for reverse in (False, True):
            edges, items, stats = emul(e, L, reverse)
            assert composite_records(edges, items) == want, (it, reverse)
            check_twins(edges)
        merged += int(stats[0])
    assert merged > 1000


@pytest.mark.parametrize("name", FIXTURES)
def test_sequential_cpp_restatement_matches_reference(name, emul):
    """oracle/contract_seq.cpp (the golden for BASELINE.json sizes) against the unmodified reference's fixtures, and its checksum
    against the numpy form of the same figure computed on what the kernel bodies return."""
    from contract_lib import seq_records, seq_simplify, simplified_checksum
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    if "c_edges" not in g.files:
        pytest.skip("fixture predates the contracted dump")
    st, edges, items = seq_simplify(g["edges"], g["len"])
    assert seq_records(edges, items) == fixture_records(g)
    ke, ki, _ = emul(g["edges"], g["len"])
    assert simplified_checksum(ke, ki) == st["checksum"]
    assert st["n_edges"] == len(ke) and st["n_items"] == len(ki)


def test_kernel_bodies_match_sequential_restatement_at_scale(emul):
    """Config 3 at a quarter (2.1 M reads, 4.25 M edges: hubs, parallel chains, thousands of ready nodes per round): the kernel bodies
    against the sequential C++ restatement, by checksum and counters. OGB_TEST_FULL_SIZE=1: the whole of config 3."""
    from contract_lib import edges_struct, seq_simplify, simplified_checksum
    from oracle_lib import LeanOracle
    from metagenomics_b200 import synth
    cfg = synth.config(3, scale=1.0 if os.environ.get("OGB_TEST_FULL_SIZE") else 0.25)
    o = LeanOracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run(keep_edges=True)
    e, L = o.edges(), o.read_info()["len"]
    st, _, _ = seq_simplify(e, L, arrays=False)
    ke, ki, ks = emul(None, L, presorted=edges_struct(e))
    assert (int(ks[0]), int(ks[1]), int(ks[2])) == (st["merges"], st["dead_ends"], st["iterations"])
    assert (len(ke), len(ki)) == (st["n_edges"], st["n_items"])
    assert simplified_checksum(ke, ki) == st["checksum"]


def test_large_random_graphs_kernel_bodies_match_sequential_restatement(emul):
    """Random multigraphs of up to 4000 nodes (long chains with random indices, hubs, parallel paths, cycles, self-loops, dead-end
    tips): the kernel bodies, threads run forwards and backwards, against the sequential C++ restatement -- records, not checksums."""
    from contract_lib import seq_records, seq_simplify, simplified_checksum
    rng = np.random.default_rng(7)
    deep = 0
    for it in range(60):
        n = int(rng.integers(200, 4000))
        e, L = random_graph(rng, n, int(rng.integers(0, n // 2)), float(rng.uniform(0.9, 0.999)))
        st, se, si = seq_simplify(e, L)
        want = seq_records(se, si)
        for reverse in (False, True):
            ke, ki, ks = emul(e, L, reverse)
            assert composite_records(ke, ki) == want, (it, reverse)
            assert simplified_checksum(ke, ki) == st["checksum"]
            assert (int(ks[0]), int(ks[1]), int(ks[2])) == (st["merges"], st["dead_ends"], st["iterations"])
        deep += st["iterations"] > 3
    assert deep > 5
