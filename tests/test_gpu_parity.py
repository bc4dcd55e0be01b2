"""GPU parity (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, must be
bit-exact with the oracle for every stage of the hot path -- K1 index content, K2 superReadID, K3
pre-reduction edges, K5/K6 post-reduction edges -- on seeded data sets and the committed goldens."""
import glob
import os

import numpy as np
import pytest

import datasets
from oracle_lib import Oracle, sort_tuples

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    from metagenomics_b200 import Context
    c = Context(0)
    yield c
    c.close()


def build_gpu(ctx, cfg, keep_pre=True):
    from metagenomics_b200 import Dataset, HashTable, OverlapGraph
    ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
    ht = HashTable(ctx)
    ht.insertDataset(ds, cfg["min_overlap"])
    og = OverlapGraph(ht, keep_pre=keep_pre)
    return ds, ht, og


def assert_same_edges(got, want, what):
    got = sort_tuples(got)
    if got.shape != want.shape or not np.array_equal(got, want):
        a = set(map(tuple, got.tolist())); b = set(map(tuple, want.tolist()))
        raise AssertionError(f"{what}: {len(got)} vs {len(want)} edges; only-gpu {sorted(a - b)[:5]} only-oracle {sorted(b - a)[:5]}")


ALL = datasets.small_configs() + datasets.adversarial()


@pytest.mark.parametrize("cfg", ALL, ids=[c["name"][:28] for c in ALL])
def test_graph_matches_oracle(ctx, cfg):
    from metagenomics_b200 import edges_as_tuples
    ds, ht, og = build_gpu(ctx, cfg)
    orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
    assert ds.getNumberOfUniqueReads() == orc.n
    # K2: contained-read flags (superReadID, Read.h:50)
    assert np.array_equal(og.superReadIDs()[1:], orc.read_info()["sup"])
    # K3: pre-reduction edge multiset (every insertEdge of the reference, OverlapGraph.cpp:416-417)
    assert_same_edges(edges_as_tuples(og.edges(pre=True)), orc.edges(pre=True), "pre-reduction")
    # K5+K6: post-reduction edge multiset, node and edge counters (OverlapGraph.cpp:394-397,638-659)
    fin = og.edges()
    assert_same_edges(edges_as_tuples(fin), orc.edges(), "post-reduction")
    c = orc.counters()
    assert og.getNumberOfEdges() == c["number_of_edges"]
    assert og.getNumberOfNodes() == c["number_of_nodes"]
    # the C ABI promises canonical order (src, offset, dst, orient)
    t = edges_as_tuples(fin)
    assert np.array_equal(t, sort_tuples(t))


@pytest.mark.parametrize("cfg", [datasets.small_configs()[0], datasets.small_configs()[4], datasets.tandem(), datasets.even_h()],
                         ids=["config1", "config5", "tandem", "even_h"])
def test_hash_table_content(ctx, cfg):
    """K1: every key's bucket = the reference's getListOfReads (id | o<<62, ascending id then o)."""
    ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
    orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"])
    h = orc.build_index()
    assert ht.getHashStringLength() == h == cfg["min_overlap"] - 1
    assert ht.getHashTableSize() > 8 * orc.n           # at least the reference's slot budget (HashTable.cpp:56)
    keys = []
    for i in range(1, orc.n + 1, max(1, orc.n // 700)):
        f, r = orc.get_read(i), orc.get_read(i, True)
        keys += [f[:h], f[-h:], r[:h], r[-h:], f[1:1 + h]]
    keys.append("ACGT" * 100)
    keys = [k[:h] for k in keys]
    got = ht.getListsOfReads(keys)
    for k, g in zip(keys, got):
        assert np.array_equal(g, orc.lookup(k)), k


def test_golden_vectors(ctx):
    """Fixtures dumped from the UNMODIFIED reference (tests/golden/make_golden.py)."""
    from metagenomics_b200 import edges_as_tuples
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert files, "no golden fixtures committed"
    for f in files:
        z = np.load(f)
        cfg = dict(bases=z["bases"], offsets=z["offsets"], min_overlap=int(z["min_overlap"]))
        ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
        assert ds.getNumberOfUniqueReads() == len(z["sup"]), f
        assert np.array_equal(og.superReadIDs()[1:], z["sup"]), f
        assert np.array_equal(ds.frequencies(), z["freq"]), f
        assert_same_edges(edges_as_tuples(og.edges()), z["edges"], os.path.basename(f))
        assert og.getNumberOfNodes() == int(z["number_of_nodes"]) and og.getNumberOfEdges() == int(z["number_of_edges"])


def test_ascii_upload_equals_packed_upload(ctx):
    """K0 both ways: ogb_reads_upload (ASCII) and ogb_reads_upload_packed give the same graph."""
    import ctypes as C
    from metagenomics_b200 import edges_as_tuples
    from metagenomics_b200._lib import check, lib
    from metagenomics_b200.api import EDGE_DTYPE
    for cfg in (datasets.small_configs()[0], datasets.small_configs()[4]):
        ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
        want = og.edges()
        orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"])
        bases, offs = orc.sorted_reads()
        check(lib().ogb_reads_upload(ctx._h, bases.ctypes.data, offs.ctypes.data, orc.n))
        check(lib().ogb_hash_build(ctx._h, cfg["min_overlap"]))
        check(lib().ogb_build_graph(ctx._h, 0))
        n = C.c_uint64()
        check(lib().ogb_graph_edge_count(ctx._h, 0, C.byref(n)))
        got = np.zeros(n.value, dtype=EDGE_DTYPE)
        check(lib().ogb_graph_edges(ctx._h, 0, got.ctypes.data, n.value))
        assert np.array_equal(edges_as_tuples(got), edges_as_tuples(want))


def test_rebuild_is_idempotent_and_symmetric(ctx):
    """Size-independent properties on a larger input (config-2 shape, ~75k reads): building twice gives
    identical output; every surviving edge has its twin (u,v,o,off) <-> (v,u,twin(o),off) for equal
    lengths (OverlapGraph.cpp:410-412); output is sorted."""
    from metagenomics_b200 import edges_as_tuples, synth
    cfg = synth.config(2, scale=0.05)
    ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
    a = edges_as_tuples(og.edges())
    og.buildOverlapGraphFromHashTable()
    b = edges_as_tuples(og.edges())
    assert np.array_equal(a, b)
    assert np.array_equal(a, sort_tuples(a))
    twin = a.copy()
    twin[:, 0], twin[:, 1] = a[:, 1], a[:, 0]
    twin[:, 3] = np.array([3, 1, 2, 0], dtype=np.uint32)[a[:, 3]]
    assert np.array_equal(sort_tuples(twin), a)
    st = ctx.stats()
    assert st["edges_final"] == len(a) and st["edges_pre"] >= len(a)
    orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.THREE_PHASE, threads=8)
    assert np.array_equal(a, orc.edges())
    c = orc.counters()
    assert st["edges_pre"] == c["E_pre"] and st["pivot_entries"] == c["T"] and st["overlap_probes"] == c["P_e"]


def test_empty_and_tiny_inputs(ctx):
    """Edge cases: no good reads; one read; two overlapping reads."""
    from metagenomics_b200 import edges_as_tuples
    for reads, m, n_edges in ((["ACGT"], 10, 0), (["ACGTTGCAAGGCTTAACCGGATATCGCGAATTC"], 10, 0)):
        cfg = datasets.from_strings(reads, m, "tiny")
        ds, ht, og = build_gpu(ctx, cfg)
        assert og.getNumberOfEdges() == n_edges and len(og.edges(pre=True)) == 0
    g = "ACGTTGCAAGGCTTAACCGGATATCGCGAATTCAGGTCCATGCAAGT"
    cfg = datasets.from_strings([g[:36], g[8:44]], 12, "pair")
    ds, ht, og = build_gpu(ctx, cfg)
    orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
    assert len(orc.edges()) == 2
    assert_same_edges(edges_as_tuples(og.edges()), orc.edges(), "pair")


def same_unitig_records(path, z, name):
    """The .unitig file of the drop-in against the unmodified reference's graph after its fix-point (the c_* arrays of the
    fixture): every record with source < destination, lists included; of a self-edge's twin pair the reference writes the one
    with the lower heap address, so there either twin is accepted."""
    from contract_lib import fixture_records, unitig_records
    got, want = unitig_records(path), fixture_records(z)
    assert [r for r in got if r[0] < r[1]] == [r for r in want if r[0] < r[1]], name
    loops_got, loops_want = [r for r in got if r[0] == r[1]], [r for r in want if r[0] == r[1]]
    assert all(r[0] <= r[1] for r in got) and 2 * len(loops_got) == len(loops_want) and set(loops_got) <= set(loops_want), name


def test_cpp_dropin_matches_reference_dump(ctx, tmp_path):
    """The C++ drop-in classes (metagenomics_b200/host: Dataset, HashTable, OverlapGraph, Edge) driven
    exactly like MetaGenomics/main.cpp:33,45-47 by host/ogb_overlap produce the reference's dump."""
    import subprocess
    from metagenomics_b200 import synth
    from oracle_lib import read_dump
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "metagenomics_b200", "host", "ogb_overlap")
    assert os.path.exists(exe), "metagenomics_b200/host/ogb_overlap not built"
    for name in ("config2_small", "config5_small", "palindromes", "tandem_mixed"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        fa, dump = str(tmp_path / (name + ".fa")), str(tmp_path / (name + ".bin"))
        synth.write_fasta(fa, z["bases"], z["offsets"])
        kind = "-pe" if name == "config2_small" else "-se"          # paired input also fills the mate-pair lists
        prefix = str(tmp_path / name)
        subprocess.run([exe, "-l", str(int(z["min_overlap"])), kind, "1", fa, "-f", prefix, "--dump", dump], check=True, timeout=120)
        d = read_dump(dump)
        # main.cpp:48-50: the sorted reads and the .unitig file -- the graph after the fix-point of OverlapGraph.cpp:211-215, run on the
        # device (--dump above is the graph at :210, then the driver finishes the constructor with simplifyGraph())
        assert os.path.exists(prefix + "_sortedReads.fasta") and os.path.exists(prefix + ".unitig")
        same_unitig_records(prefix + ".unitig", z, name)
        # without --dump the constructor runs the fix-point itself: same file
        subprocess.run([exe, "-l", str(int(z["min_overlap"])), kind, "1", fa, "-f", prefix + "_b"], check=True, timeout=120)
        assert open(prefix + "_b.unitig", "rb").read() == open(prefix + ".unitig", "rb").read(), name
        # the resume path (main.cpp:36-42) reads the file back: both directions of every composite edge
        subprocess.run([exe, "-l", str(int(z["min_overlap"])), kind, "1", fa, "-f", prefix, "-s", "--dump", dump + "2"], check=True, timeout=120)
        d2 = read_dump(dump + "2")
        ce = z["c_edges"].astype(np.int64)
        assert np.array_equal(d2["edges"], sort_tuples(np.stack([ce[:, 0], ce[:, 1], ce[:, 4], ce[:, 2]], axis=1))), name
        assert d2["number_of_edges"] == int(z["c_number_of_edges"]) and d2["number_of_nodes"] == int(z["c_number_of_nodes"]), name
        assert d["n"] == len(z["sup"]) and np.array_equal(d["reads"]["fnv"], z["fnv"]), name
        assert np.array_equal(d["reads"]["sup"], z["sup"]) and np.array_equal(d["reads"]["freq"], z["freq"]), name
        assert np.array_equal(d["edges"], z["edges"]), name
        assert d["number_of_nodes"] == int(z["number_of_nodes"]) and d["number_of_edges"] == int(z["number_of_edges"]), name


@pytest.mark.parametrize("cap", [8, 16])
def test_heavy_lists_match_oracle(ctx, cap, monkeypatch):
    """Slot regions of only `cap` edge words (OGB_SLOT_CAP): most nodes become heavy nodes whose lists are
    gathered in the extension area (k_heavy_move / k_heavy_place), K5 takes the any-degree path for degree
    > 32 and K6 / k_emit read the moved lists -- same graph as the oracle."""
    from metagenomics_b200 import edges_as_tuples, synth
    monkeypatch.setenv("OGB_SLOT_CAP", str(cap))
    for cfg in (synth.config(2, scale=0.02), datasets.tandem(), synth.containment_stress(7, genome_len=6000, n_primary=1500)):
        ds, ht, og = build_gpu(ctx, cfg)
        orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
        assert ctx.stats()["overflow_reads"] > 0 or orc.counters()["max_degree"] <= cap
        assert_same_edges(edges_as_tuples(og.edges(pre=True)), orc.edges(pre=True), f"pre-reduction, cap {cap}")
        assert_same_edges(edges_as_tuples(og.edges()), orc.edges(), f"post-reduction, cap {cap}")
        assert og.getNumberOfNodes() == orc.counters()["number_of_nodes"]


DS_SETS = datasets.small_configs() + datasets.adversarial()


@pytest.mark.parametrize("cfg", DS_SETS, ids=[c["name"][:28] for c in DS_SETS])
def test_device_dataset_equals_host_dataset(ctx, cfg):
    """Dataset stage on the GPU (canonical strand, LSD radix sort over packed words, dedupe, frequencies) against the
    host implementation, which the CPU suite pins to the oracle and the reference: same IDs, lengths, frequencies and
    packed words; and the graph built from the reads it leaves in HBM is the oracle's."""
    from metagenomics_b200 import Dataset, HashTable, OverlapGraph, edges_as_tuples
    host = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
    dev = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"], device=ctx)
    assert dev.getNumberOfReads() == host.getNumberOfReads() and dev.getNumberOfUniqueReads() == host.getNumberOfUniqueReads()
    assert dev.shortestReadLength == host.shortestReadLength and dev.longestReadLength == host.longestReadLength
    assert np.array_equal(dev.lengths(), host.lengths()) and np.array_equal(dev.frequencies(), host.frequencies())
    hw, ho, _ = host.packed(); dw, do, _ = dev.packed()
    assert np.array_equal(do, ho) and np.array_equal(dw, hw)
    ht = HashTable(ctx)
    ht.insertDataset(dev, cfg["min_overlap"])          # nothing to upload: the reads are resident
    og = OverlapGraph(ht)
    orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
    assert_same_edges(edges_as_tuples(og.edges()), orc.edges(), "post-reduction, device Dataset")
    assert np.array_equal(og.superReadIDs()[1:], orc.read_info()["sup"])


def test_device_dataset_config2_scale(ctx):
    from metagenomics_b200 import Dataset, synth
    cfg = synth.config(2, scale=0.2)
    host = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
    dev = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"], device=ctx)
    assert np.array_equal(dev.frequencies(), host.frequencies())
    hw, ho, _ = host.packed(); dw, do, _ = dev.packed()
    assert np.array_equal(do, ho) and np.array_equal(dw, hw)


def test_partitioned_probe_with_skewed_partitions(ctx, monkeypatch):
    """Eight hash partitions forced on small, repetitive data sets (OGB_SUB_PARTITIONS): the windows of a tandem array or
    an inverted repeat share their leading bases, so one partition's window queue receives far more than its even
    share -- the build retries with more slack and still returns the oracle's graph."""
    from metagenomics_b200 import edges_as_tuples, synth
    monkeypatch.setenv("OGB_SUB_PARTITIONS", "8")
    for cfg in (datasets.tandem(), datasets.repeats(), datasets.palindromes(), datasets.from_strings(["ACACACACACACACACACACACACACACACACACACACACACAC" + "G" * i + "T" for i in range(1, 12)], 10, "same lead"), synth.config(2, scale=0.02),
                datasets.tandem(mixed=True), synth.containment_stress(7, genome_len=6000, n_primary=1500)):   # mixed lengths: K2 through the queues too
        ds, ht, og = build_gpu(ctx, cfg)
        orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
        assert np.array_equal(og.superReadIDs()[1:], orc.read_info()["sup"])
        assert_same_edges(edges_as_tuples(og.edges(pre=True)), orc.edges(pre=True), "pre-reduction, 8 partitions")
        assert_same_edges(edges_as_tuples(og.edges()), orc.edges(), "post-reduction, 8 partitions")


def test_skewed_keys_fill_a_hash_partition(ctx, monkeypatch):
    """K1 bound: with 16 partitions forced, a quarter of all keys -- the primer in front of every read
    -- lands in one partition that holds 10/16 slots per read: the insert must give up after one lap (CTR_TABLE_FULL) and
    ogb_hash_build must retry with fewer partitions instead of spinning; the result is still the oracle's graph, and the
    index still answers getListOfReads like the reference's (HashTable.cpp:202-221)."""
    from metagenomics_b200 import edges_as_tuples
    monkeypatch.setenv("OGB_SUB_PARTITIONS", "16")
    cfg = datasets.primer_prefixed()
    ds, ht, og = build_gpu(ctx, cfg)
    orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.THREE_PHASE, threads=8)
    assert ds.getNumberOfUniqueReads() == orc.n and orc.n > 20000 and orc.counters()["E_pre"] > 100000
    st = ctx.stats()
    assert st["hash_build_attempts"] >= 2 and st["hash_partitions"] < 16  # one partition of 16 could not hold the primer keys: K1 gave up and was retried
    assert_same_edges(edges_as_tuples(og.edges(pre=True)), orc.edges(pre=True), "pre-reduction, primer-prefixed")
    assert_same_edges(edges_as_tuples(og.edges()), orc.edges(), "post-reduction, primer-prefixed")
    h = orc.build_index()
    keys = [orc.get_read(i)[:h] for i in range(1, orc.n + 1, orc.n // 50)] + ["GATTACAGGCCTTAGCAATC" + "A" * (h - 20)]
    for k, g in zip(keys, ht.getListsOfReads(keys)):
        assert np.array_equal(g, orc.lookup(k)), k


@pytest.mark.parametrize("name", ["config2_small", "paired_mixed"])
@pytest.mark.parametrize("where", ["device", "host"])
def test_mate_pair_lists_match_reference(ctx, tmp_path, name, where):
    """Dataset::storeMatePairInformation (Dataset.cpp:208-310, called at OverlapGraph.cpp:142): every read's mate-pair list --
    mate ID after super-read redirection, the two orientation bits, data set number, in list order -- as the UNMODIFIED reference
    leaves it (fixture dumped through oracle/ref_harness.cpp --mates). `device`: the batched lookup kernel (ogb_mate_lookup:
    filter, getReadFromString as one verified index lookup, redirection, substring test); `host`: the per-pair loop of the
    drop-in class. paired_mixed has 81 % contained reads, i.e. most mates are redirected."""
    import subprocess
    from metagenomics_b200 import synth
    from oracle_lib import read_dump, read_mates
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "metagenomics_b200", "host", "ogb_overlap")
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fa, dump, mates = str(tmp_path / "in.fa"), str(tmp_path / "d.bin"), str(tmp_path / "m.bin")
    synth.write_fasta(fa, z["bases"], z["offsets"])
    env = dict(os.environ, **({"OGB_MATES_ON_HOST": "1"} if where == "host" else {}))
    subprocess.run([exe, "-l", str(int(z["min_overlap"])), "-pe", "1", fa, "--dump", dump, "--mates", mates], check=True, timeout=120, env=env)
    d = read_dump(dump)
    assert np.array_equal(d["reads"]["sup"], z["sup"]) and np.array_equal(d["edges"], z["edges"])
    start, lists = read_mates(mates)
    assert np.array_equal(start, z["mate_start"]) and np.array_equal(lists, z["mate_lists"])
    assert len(lists) > 1000


def test_mate_lookup_entry_point(ctx):
    """ogb_mate_lookup directly: bad sequences (N, low complexity, too short) give 0; a sequence and its reverse complement find
    the same read with opposite orientation bits."""
    import ctypes as C
    from metagenomics_b200._lib import check, lib
    cfg = datasets.small_configs()[1]
    ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
    first = ds.getReadFromID(7)
    seqs = [first, datasets.rc(first), "N" + first[1:], "A" * len(first), first[:cfg["min_overlap"]]]
    flat = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    ids, ori = np.zeros(len(seqs), np.uint32), np.zeros(len(seqs), np.uint8)
    check(lib().ogb_mate_lookup(ctx._h, flat.ctypes.data, offs.ctypes.data, len(seqs), cfg["min_overlap"], ids.ctypes.data, ori.ctypes.data))
    assert ids.tolist() == [7, 7, 0, 0, 0]
    assert ori[0] == 1 and (ori[1] == 0 or first == datasets.rc(first))


# ---- the simplification stage (OverlapGraph.cpp:211-215) ----

def test_simplify_matches_reference_fixtures(ctx):
    """contractCompositePaths + removeDeadEndNodes to the fix-point, on the device, against the unmodified reference's graph after
    that stage (--dump2 fixtures): end points, orientation, offset and the read / offset / orientation lists of every edge."""
    from contract_lib import check_twins, composite_records, fixture_records
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    seen = 0
    for f in files:
        z = np.load(f)
        if "c_edges" not in z.files:
            continue
        cfg = dict(bases=z["bases"], offsets=z["offsets"], min_overlap=int(z["min_overlap"]))
        ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
        edges, items, st = og.simplify()
        assert composite_records(edges, items) == fixture_records(z), os.path.basename(f)
        check_twins(edges)
        assert st["n_edges_in"] == int(z["number_of_edges"]) and st["n_edges_out"] == len(z["c_edges"]) == int(z["c_number_of_edges"]), f
        assert len(np.unique(edges["src"])) == int(z["c_number_of_nodes"]), f
        # the graph of ogb_graph_edges is untouched, and a second run gives the same result
        e2, i2, st2 = og.simplify()
        assert np.array_equal(e2, edges) and np.array_equal(i2, items) and og.getNumberOfEdges() == int(z["number_of_edges"])
        seen += 1
    assert seen >= 6


def test_simplify_seeded_sets_match_sequential_restatement(ctx):
    """Beyond the fixtures: the adversarial sets (parallel chains between the same end nodes, tandem repeats, palindromes ...)
    and samples of configs 1-5 against the sequential restatement (pinned to the reference by the CPU suite)."""
    from contract_lib import check_twins, composite_records, load_oracle_module
    from metagenomics_b200 import edges_as_tuples, synth
    co = load_oracle_module("contract_oracle")
    sets = datasets.adversarial() + [synth.config(2, scale=0.02), synth.config(3, scale=0.004), synth.config(1, scale=1.0),
                                     synth.containment_stress(9, genome_len=9000, n_primary=2500), synth.config(4, scale=0.001), datasets.paired_mixed()]
    for cfg in sets:
        ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
        t = edges_as_tuples(og.edges())
        want = co.Graph(t.tolist(), ds.lengths().tolist()).simplify().edge_records()
        edges, items, st = og.simplify()
        assert composite_records(edges, items) == want, cfg["name"]
        check_twins(edges)


def test_simplify_config3_quarter_equals_kernel_bodies_on_cpu(ctx):
    """2.1 M reads / 4.2 M edges: the launches on the device (thousands of nodes per round, races would show here) against the
    same per-thread bodies run one after the other on the CPU (tests/contract_emul.cpp; the bodies are pinned to the reference
    by tests/test_contract.py). Same deterministic output order, so the arrays must be identical. Plus the invariants of the
    stage: a contracted read sits in exactly one edge pair and owns no edge, offsets add up."""
    from contract_lib import build_emul, check_twins
    from metagenomics_b200 import synth
    cfg = synth.config(3, scale=0.25)
    ds, ht, og = build_gpu(ctx, cfg, keep_pre=False)
    fin = og.edges()
    edges, items, st = og.simplify()
    want_e, want_i, stats = build_emul()(None, ds.lengths(), presorted=fin)
    assert st["merges"] == int(stats[0]) and st["dead_ends"] == int(stats[1]) and st["iterations"] == int(stats[2])
    assert np.array_equal(edges, want_e) and np.array_equal(items, want_i)
    check_twins(edges)
    inside = items["read"]
    reads, counts = np.unique(inside, return_counts=True)
    assert (counts == 2).all()                                                  # once per direction
    assert not np.isin(reads, edges["src"]).any()
    sums = np.add.reduceat(items["offset"].astype(np.int64), edges["list_start"][edges["count"] > 0].astype(np.int64)) if len(items) else np.array([])
    assert (sums <= edges["offset"][edges["count"] > 0].astype(np.int64)).all()
    print("simplify config3@0.25:", st)
