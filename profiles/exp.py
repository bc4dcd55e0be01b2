"""Torch-free experiment runner (starts in seconds on a fresh box): builds one configuration K times through the
C ABI, prints the per-phase device times and checks the final edge set against tests/golden/full_size.json.

    python profiles/exp.py --config 2 --scale 1.0 --steps 5 --warmup 2 [--tag name]
Environment knobs of libogb (OGB_*) apply; OGB_LIB=<path> loads another build of the library."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def checksum(e):
    e = e.astype(np.uint64)
    x = (e[:, 0] * np.uint64(0x9E3779B97F4A7C15) ^ e[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) ^ e[:, 2] * np.uint64(0x165667B19E3779F9)
         ^ e[:, 3] * np.uint64(0x27D4EB2F165667C5))
    x ^= x >> np.uint64(29); x *= np.uint64(0xBF58476D1CE4E5B9); x ^= x >> np.uint64(32)
    return [int(np.bitwise_xor.reduce(x)), int(x.sum(dtype=np.uint64))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--tag", default="")
    ap.add_argument("--device-dataset", action="store_true")
    ap.add_argument("--simplify", action="store_true", help="also run the simplification stage (OverlapGraph.cpp:211-215) on the last graph")
    a = ap.parse_args()
    import ctypes as C
    from metagenomics_b200 import Context, Dataset, HashTable, OverlapGraph, edges_as_tuples, synth
    from metagenomics_b200._lib import check, lib
    cfg = synth.config(a.config, scale=a.scale)
    ctx = Context(0)
    t0 = time.time()
    ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
    t_host = time.time() - t0
    if a.device_dataset:
        for rep in range(2):                       # second run: pools and CUDA modules are warm
            t0 = time.time()
            dd = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"], device=ctx)
            t_dev = time.time() - t0
        same = (np.array_equal(dd.frequencies(), ds.frequencies()) and np.array_equal(dd.packed()[0], ds.packed()[0]))
        print(f"[{a.tag or 'exp'}] Dataset stage: host threads {t_host:.3f} s, device {t_dev:.3f} s, identical: {same}", flush=True)
        ds = dd
    ht = HashTable(ctx)
    ht.insertDataset(ds, cfg["min_overlap"])
    og = OverlapGraph(ht)
    gold = [g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "full_size.json"))) if g["config"] == a.config and g["scale"] == a.scale]
    verdict = "no golden"
    if gold:
        verdict = "PARITY OK" if checksum(edges_as_tuples(og.edges())) == gold[0]["checksum"] else "PARITY MISMATCH"
    rows = []
    for i in range(a.warmup + a.steps):
        check(lib().ogb_l2_flush(ctx._h, 512 << 20))
        check(lib().ogb_timer_begin(ctx._h))
        check(lib().ogb_hash_build(ctx._h, cfg["min_overlap"]))
        check(lib().ogb_mark_contained(ctx._h))
        check(lib().ogb_build_graph(ctx._h, 0))
        ms = C.c_float()
        check(lib().ogb_timer_end(ctx._h, C.byref(ms)))
        if i >= a.warmup:
            st = ctx.stats()
            rows.append([ms.value] + [st[k] for k in ("ms_hash_build", "ms_contain", "ms_overlap", "ms_scan_kernel", "ms_probe_launch", "ms_exchange_pre", "ms_mark", "ms_reduce", "ms_window_launch")])
    r = np.array(rows).mean(axis=0)
    st = ctx.stats()
    print(f"[{a.tag or 'exp'}] config{a.config}@{a.scale} n={st['n_reads']} {verdict} | step {r[0]:.3f} ms | hash {r[1]:.3f} contain {r[2]:.3f} overlap {r[3]:.3f} "
          f"(scan {r[4]:.3f}, probe launch {r[5]:.4f} [window part {r[9]:.4f}] x{st['probe_launches']}) barrier {r[6]:.3f} mark {r[7]:.3f} reduce {r[8]:.3f} | "
          f"E_pre {st['edges_pre']} E_final {st['edges_final']} heavy {st['overflow_reads']} launches {st['kernel_launches']} | host setup {t_host:.1f} s", flush=True)
    if a.simplify:
        for _ in range(2):
            edges, items, ss = og.simplify()
        print(f"[{a.tag or 'exp'}] simplify: {ss}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
