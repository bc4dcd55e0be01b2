"""GPU, BASELINE.json sizes: the build at full configuration sizes, checked against golden counters
and an order-independent checksum produced by the CPU oracle (three-phase form, 8 threads, run once in
the build container: tests/golden/full_size.json) and through size-independent properties (twin
symmetry, canonical order, idempotence); and the simplification stage that follows (OverlapGraph.cpp:211-215) against
the golden of its sequential restatement."""
import json
import os

import numpy as np
import pytest

from oracle_lib import sort_tuples

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
# one GPU: the entries up to ~10 M reads (the larger ones -- config 3 x 2/4/8, config 4 at full size -- are the multi-GPU cases of
# tests/test_multi_gpu.py and the bench's own parity check)
GOLD = [g for g in json.load(open(os.path.join(HERE, "golden", "full_size.json"))) if g["n_unique"] <= 11_000_000 and (g["config"], g["scale"]) != (2, 4.0)]


def checksum(e):
    e = e.astype(np.uint64)
    x = (e[:, 0] * np.uint64(0x9E3779B97F4A7C15) ^ e[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) ^ e[:, 2] * np.uint64(0x165667B19E3779F9)
         ^ e[:, 3] * np.uint64(0x27D4EB2F165667C5))
    x ^= x >> np.uint64(29); x *= np.uint64(0xBF58476D1CE4E5B9); x ^= x >> np.uint64(32)
    return [int(np.bitwise_xor.reduce(x)), int(x.sum(dtype=np.uint64))]


@pytest.mark.parametrize("gold", GOLD, ids=[f"config{g['config']}@{g['scale']}" for g in GOLD])
def test_full_size_matches_oracle_checksum(gold):
    from metagenomics_b200 import Context, Dataset, HashTable, OverlapGraph, edges_as_tuples, synth
    cfg = synth.config(gold["config"], scale=gold["scale"])
    ctx = Context(0)
    try:
        ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
        assert ds.getNumberOfUniqueReads() == gold["n_unique"]
        ht = HashTable(ctx)
        ht.insertDataset(ds, cfg["min_overlap"])
        og = OverlapGraph(ht)
        st = ctx.stats()
        assert st["edges_pre"] == gold["E_pre"] and st["edges_final"] == gold["E_final"] and st["nodes_final"] == gold["nodes"]
        assert st["n_contained"] == gold["contained"] and st["max_degree"] == gold["max_degree"]
        assert st["pivot_entries"] == gold["T"]
        t = edges_as_tuples(og.edges())
        assert checksum(t) == gold["checksum"]
        # properties: canonical order; every edge has its twin (offsets equal for equal lengths, else shifted by the length difference)
        assert np.array_equal(t, sort_tuples(t)) if len(t) < 5_000_000 else bool(np.all(np.diff(t[:, 0].astype(np.int64)) >= 0))
        lens = ds.lengths().astype(np.int64)
        tw = t.copy()
        tw[:, 0], tw[:, 1] = t[:, 1], t[:, 0]
        tw[:, 3] = np.array([3, 1, 2, 0], dtype=np.uint32)[t[:, 3]]
        tw[:, 2] = ((lens[t[:, 1] - 1] + t[:, 2].astype(np.int64) - lens[t[:, 0] - 1]) & 0xFFFF).astype(np.uint32)
        assert checksum(tw) == gold["checksum"]
        # the simplification stage (OverlapGraph.cpp:211-215) against the sequential restatement's golden: counters of the fix-point and
        # the checksum over every composite edge with its read / offset / orientation lists (tests/golden/make_simplified_full_size.py)
        if "simplified" in gold:
            from contract_lib import check_twins, simplified_checksum
            want = gold["simplified"]
            edges, items, sst = og.simplify()
            assert (sst["n_edges_out"], sst["n_items"], sst["merges"], sst["dead_ends"], sst["iterations"]) == \
                (want["n_edges"], want["n_items"], want["merges"], want["dead_ends"], want["iterations"])
            assert simplified_checksum(edges, items) == want["checksum"]
            check_twins(edges)
        og.buildOverlapGraphFromHashTable()                    # idempotence
        assert checksum(edges_as_tuples(og.edges())) == gold["checksum"]
    finally:
        ctx.close()
