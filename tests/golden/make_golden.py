"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/ref_overlap,
built by oracle/Makefile from /root/reference/MetaGenomics). Run in the build container:

    python tests/golden/make_golden.py

Each .npz holds the raw input reads and the reference's own output at OverlapGraph.cpp:210:
edges (src, dst, overlapOffset, orientation) sorted canonically, superReadID, frequency, read
lengths, fnv1a of every forward string (pins the Dataset sort/dedupe order), numberOfNodes/Edges -- and, for the next row of
SURVEY.md 8(f), the reference graph after its contractCompositePaths / removeDeadEndNodes fix-point (:211-215):
c_edges (src, dst, orient, nlist, overlapOffset), c_list_start, c_lists (read, overlapOffset, orientation)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import datasets  # noqa: E402
from metagenomics_b200 import synth  # noqa: E402
from oracle_lib import have_reference, run_reference  # noqa: E402


def main():
    assert have_reference(), "build oracle/_ref first (make -C oracle)"
    sets = {
        "config1_small": synth.config(1, scale=0.08),
        "config2_small": synth.config(2, scale=0.0012),
        "config4_small": synth.config(4, scale=0.00008),
        "config5_small": synth.containment_stress(5, genome_len=6000, n_primary=1500),
        "tandem_mixed": datasets.tandem(mixed=True),
        "repeats": datasets.repeats(),
        "palindromes": datasets.palindromes(),
        "filtered": datasets.filtered(),
        "paired_mixed": datasets.paired_mixed(),
    }
    for name, cfg in sets.items():
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "in.fa")
            synth.write_fasta(fa, cfg["bases"], cfg["offsets"])
            # the .unitig file the unmodified reference writes for this graph (small text; fixture of the reader / writer test)
            d, t, _ = run_reference([fa], cfg["min_overlap"], paired=cfg["paired"], contracted=True, unitig=os.path.join(HERE, name + ".unitig"), mates=True)
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, bases=cfg["bases"], offsets=cfg["offsets"], min_overlap=np.int64(cfg["min_overlap"]),
                            edges=d["edges"], sup=d["reads"]["sup"], freq=d["reads"]["freq"], len=d["reads"]["len"],
                            fnv=d["reads"]["fnv"], number_of_nodes=np.int64(d["number_of_nodes"]),
                            number_of_edges=np.int64(d["number_of_edges"]), n_good=np.int64(t["n_reads"]),
                            c_edges=d["contracted"]["edges"], c_list_start=d["contracted"]["list_start"], c_lists=d["contracted"]["lists"],
                            c_number_of_nodes=np.int64(d["contracted"]["number_of_nodes"]), c_number_of_edges=np.int64(d["contracted"]["number_of_edges"]),
                            mate_start=d["mates"][0], mate_lists=d["mates"][1], paired=np.int64(1 if cfg["paired"] else 0))
        print(f"{name}: {d['n']} unique reads, {len(d['edges'])} edges, {len(d['contracted']['edges'])} after contraction -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
