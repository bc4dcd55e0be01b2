"""TEST INFRASTRUCTURE: regenerates tests/golden/full_size.json -- counters and an order-independent
checksum of the final edge set of the CPU oracle (oracle/liboracle.so, three-phase form on all host
threads) at BASELINE.json sizes. Run once in the build container; the GPU parity test
tests/test_gpu_full_size.py compares the CUDA path with these numbers.

    python tests/golden/make_full_size.py                  # every entry of CASES
    python tests/golden/make_full_size.py 5:1.0 4:0.2      # only these config:scale pairs (merged into the file)

The three-phase form is pinned to the statement-for-statement BFS form and to the unmodified reference
on the small fixtures (tests/test_oracle_golden.py); it is the only form that finishes at these sizes.

    python tests/golden/make_full_size.py --lean 3:8.0 4:1.0   # the memory-lean form (oracle/lean_oracle.cpp)

--lean: entries the std::string port cannot hold in this container's 62 GB -- BASELINE.json configs[3] at full size
(50 M x 150 bp) and the weak-scaled bench workloads (config 3 / config 2 x 2, 4, 8). The lean form is pinned to the port
and to the reference dumps by tests/test_oracle_golden.py::test_lean_oracle_*, and reproduces the port's entries of this
file bit for bit (configs 2 and 5 at 1.0, checked when it was introduced).
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

from oracle_lib import LeanOracle, Oracle  # noqa: E402
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


def one_lean(k, scale, threads):
    t0 = time.time()
    cfg = synth.config(k, scale=scale)
    t1 = time.time()
    orc = LeanOracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"], threads=threads)
    del cfg
    t2 = time.time()
    orc.run()
    c = orc.counters()
    assert c["asymmetric"] == 0
    rec = dict(config=k, scale=scale, n_unique=c["n_unique"], E_pre=c["E_pre"], E_final=c["E_final"], nodes=c["nodes"], contained=c["contained"],
               checksum=orc.checksum(), max_degree=c["max_degree"], P_e=c["P_e"], T=c["T"], P_c=c["P_c"], C_c=c["C_c"], oracle="lean")
    print(f"config {k} @ {scale}: {rec}  (synth {t1 - t0:.0f} s, dataset stage {t2 - t1:.0f} s, build {time.time() - t2:.0f} s)", flush=True)
    return rec


def main():
    args = sys.argv[1:]
    lean = "--lean" in args
    args = [a for a in args if a != "--lean"]
    cases = [(int(a.split(":")[0]), float(a.split(":")[1])) for a in args] or CASES
    threads = os.cpu_count() or 1
    for k, s in cases:
        rec = one_lean(k, s, threads) if lean else one(k, s, threads)
        have = json.load(open(OUT)) if os.path.exists(OUT) else []
        have = [g for g in have if not (g["config"] == k and g["scale"] == s)] + [rec]
        have.sort(key=lambda g: (g["config"], g["scale"]))
        with open(OUT, "w") as f:
            json.dump(have, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
