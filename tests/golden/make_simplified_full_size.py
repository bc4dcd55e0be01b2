"""TEST INFRASTRUCTURE: adds the golden of the SIMPLIFICATION stage (OverlapGraph.cpp:211-215: contractCompositePaths +
removeDeadEndNodes to the fix-point) to entries of tests/golden/full_size.json -- the field

    "simplified": {"n_edges", "n_items", "merges", "dead_ends", "iterations", "checksum": [xor, sum]}

from oracle/contract_seq.cpp (the reference's sweep in its own sequential order; pinned to the reference's --dump2 fixtures by
tests/test_contract.py) run on the final edge list of the memory-lean oracle, whose checksum must equal the entry's.

    python tests/golden/make_simplified_full_size.py 3:1.0 2:1.0 5:1.0 4:0.2
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from contract_lib import seq_simplify  # noqa: E402
from oracle_lib import LeanOracle, edge_checksum  # noqa: E402
from metagenomics_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "full_size.json")


def main():
    want = [(int(a.split(":")[0]), float(a.split(":")[1])) for a in sys.argv[1:]]
    gold = json.load(open(OUT))
    for k, scale in want:
        entry = next(g for g in gold if g["config"] == k and g["scale"] == scale)
        t0 = time.time()
        cfg = synth.config(k, scale=scale)
        o = LeanOracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run(keep_edges=True)
        e, lens = o.edges(), o.read_info()["len"]
        assert edge_checksum(e) == entry["checksum"], "the lean oracle's final edges are not the entry's"
        st, _, _ = seq_simplify(e, lens, arrays=False)
        entry["simplified"] = st
        print(f"config {k} @ {scale}: {len(e)} edges -> {st}  ({time.time() - t0:.0f} s)", flush=True)
        json.dump(gold, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
