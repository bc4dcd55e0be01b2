"""TEST INFRASTRUCTURE: regenerates tests/golden/full_size.json -- counters and an order-independent
checksum of the final edge set of the CPU oracle (oracle/liboracle.so, three-phase form on all host
threads) at BASELINE.json sizes. Run once in the build container; the GPU parity test
tests/test_gpu_full_size.py compares the CUDA path with these numbers.

    python tests/golden/make_full_size.py                  # every entry of CASES
    python tests/golden/make_full_size.py 5:1.0 4:0.2      # only these config:scale pairs (merged into the file)

The three-phase form is pinned to the statement-for-statement BFS form and to the unmodified reference
on the small fixtures (tests/test_oracle_golden.py); it is the only form that finishes at these sizes.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import Oracle  # noqa: E402
from metagenomics_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "full_size.json")
# (config, scale): 2, 3 and 5 at BASELINE.json size; 4 at the largest scale the oracle holds in this container's RAM
CASES = [(2, 1.0), (3, 1.0), (5, 0.25), (5, 1.0), (4, 0.04), (4, 0.2)]


def checksum(e):
    """xor and sum of a 64-bit mix of every (src, dst, offset, orient) tuple: independent of order."""
    e = e.astype(np.uint64)
    x = (e[:, 0] * np.uint64(0x9E3779B97F4A7C15) ^ e[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) ^ e[:, 2] * np.uint64(0x165667B19E3779F9)
         ^ e[:, 3] * np.uint64(0x27D4EB2F165667C5))
    x ^= x >> np.uint64(29); x *= np.uint64(0xBF58476D1CE4E5B9); x ^= x >> np.uint64(32)
    return [int(np.bitwise_xor.reduce(x)), int(x.sum(dtype=np.uint64))]


def one(k, scale, threads):
    t0 = time.time()
    cfg = synth.config(k, scale=scale)
    orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"])
    del cfg
    orc.run_all(Oracle.THREE_PHASE, threads=threads)
    c = orc.counters()
    e = orc.edges()
    sup = orc.read_info()["sup"]
    rec = dict(config=k, scale=scale, n_unique=int(orc.n), E_pre=c["E_pre"], E_final=int(len(e)), nodes=c["number_of_nodes"],
               contained=int(np.count_nonzero(sup)), checksum=checksum(e), max_degree=c["max_degree"], P_e=c["P_e"], T=c["T"])
    assert c["number_of_edges"] == len(e)
    print(f"config {k} @ {scale}: {rec}  ({time.time() - t0:.0f} s)", flush=True)
    return rec


def main():
    cases = [(int(a.split(":")[0]), float(a.split(":")[1])) for a in sys.argv[1:]] or CASES
    have = json.load(open(OUT)) if os.path.exists(OUT) else []
    threads = os.cpu_count() or 1
    for k, s in cases:
        rec = one(k, s, threads)
        have = [g for g in have if not (g["config"] == k and g["scale"] == s)] + [rec]
        have.sort(key=lambda g: (g["config"], g["scale"]))
        with open(OUT, "w") as f:
            json.dump(have, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
