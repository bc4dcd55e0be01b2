"""Worker of tests/test_multi_gpu.py: every rank builds the graph of a few small data sets through the
multi-rank C ABI (NCCL inside libogb) and checks the replicated result against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import datasets  # noqa: E402
from oracle_lib import Oracle, sort_tuples  # noqa: E402
from metagenomics_b200 import Dataset, HashTable, OverlapGraph, edges_as_tuples, synth  # noqa: E402
from metagenomics_b200.dist import make_context  # noqa: E402
from contract_lib import check_twins, composite_records, load_oracle_module  # noqa: E402


def main():
    ctx, rank, world, local = make_context()
    torch.cuda.set_device(local)
    sets = [synth.config(1, scale=0.3), synth.containment_stress(5, genome_len=20000, n_primary=5000), datasets.tandem(),
            datasets.palindromes(), synth.config(2, scale=0.02), datasets.from_strings(["ACGT"], 10, "empty")]
    co = load_oracle_module("contract_oracle")
    for cfg in sets:
        ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
        ht = HashTable(ctx)
        ht.insertDataset(ds, cfg["min_overlap"])
        og = OverlapGraph(ht, keep_pre=True)
        orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
        got, pre = sort_tuples(edges_as_tuples(og.edges())), sort_tuples(edges_as_tuples(og.edges(pre=True)))
        assert np.array_equal(og.superReadIDs()[1:], orc.read_info()["sup"]), (rank, cfg["name"], "superReadID")
        assert np.array_equal(pre, orc.edges(pre=True)), (rank, cfg["name"], "pre-reduction", pre.shape, orc.edges(pre=True).shape)
        assert np.array_equal(got, orc.edges()), (rank, cfg["name"], "post-reduction", got.shape, orc.edges().shape)
        c = orc.counters()
        assert og.getNumberOfEdges() == c["number_of_edges"] and og.getNumberOfNodes() == c["number_of_nodes"], (rank, cfg["name"])
        st = ctx.stats()
        assert st["edges_pre"] == c["E_pre"] and (world == 1 or st["edges_pre_local"] <= st["edges_pre"])
        # the simplification stage (OverlapGraph.cpp:211-215) is rank-local: every rank holds the whole graph and gets the same result
        if len(got) <= 60000:
            want = co.Graph(got.tolist(), ds.lengths().tolist()).simplify().edge_records()
            edges, items, _ = og.simplify()
            assert composite_records(edges, items) == want, (rank, cfg["name"], "simplified graph")
            check_twins(edges)
    print(f"rank {rank}/{world}: {len(sets)} data sets identical to the oracle", flush=True)
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
