"""Torch-free multi-rank experiment runner: spawns one process per GPU, each builds its weak-scaled config-2 workload
through the multi-rank C ABI (NCCL + CUDA IPC inside libogb) and prints its own per-phase device times.

    python profiles/exp_mp.py --gpus 4 [--steps 4 --warmup 2 --scale-per-gpu 1.0]
The NCCL unique id is made by the parent and handed to the ranks through the environment."""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    rank, world = int(os.environ["OGB_RANK"]), int(os.environ["OGB_WORLD"])
    steps, warmup, scale = int(os.environ["OGB_STEPS"]), int(os.environ["OGB_WARMUP"]), float(os.environ["OGB_SCALE"])
    from metagenomics_b200 import Context, Dataset, HashTable, OverlapGraph, synth
    from metagenomics_b200._lib import check, lib
    cfg = synth.config(2, scale=scale * world)
    ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
    ctx = Context(rank, rank, world, bytes.fromhex(os.environ["OGB_UID"]))
    ht = HashTable(ctx)
    ht.insertDataset(ds, cfg["min_overlap"])
    og = OverlapGraph(ht)
    rows = []
    for i in range(warmup + steps):
        check(lib().ogb_l2_flush(ctx._h, 512 << 20))
        check(lib().ogb_timer_begin(ctx._h))
        check(lib().ogb_hash_build(ctx._h, cfg["min_overlap"]))
        check(lib().ogb_mark_contained(ctx._h))
        check(lib().ogb_build_graph(ctx._h, 0))
        ms = C.c_float()
        check(lib().ogb_timer_end(ctx._h, C.byref(ms)))
        if i >= warmup:
            st = ctx.stats()
            rows.append([ms.value] + [st[k] for k in ("ms_hash_build", "ms_overlap", "ms_scan_kernel", "ms_exchange_pre", "ms_mark", "ms_reduce")])
    r = np.array(rows).mean(axis=0)
    st = ctx.stats()
    print(f"[rank {rank}/{world}] n={st['n_reads']} step {r[0]:.3f} ms | hash {r[1]:.3f} overlap {r[2]:.3f} (scan {r[3]:.3f}) barrier {r[4]:.3f} "
          f"mark {r[5]:.3f} reduce {r[6]:.3f} | E_pre {st['edges_pre']} local {st['edges_pre_local']} E_final {st['edges_final']}", flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scale-per-gpu", type=float, default=1.0)
    a = ap.parse_args()
    from metagenomics_b200 import nccl_unique_id
    env = dict(os.environ, OGB_UID=nccl_unique_id().hex(), OGB_WORLD=str(a.gpus), OGB_STEPS=str(a.steps), OGB_WARMUP=str(a.warmup),
               OGB_SCALE=str(a.scale_per_gpu), OGB_WORKER="1")
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__)], env=dict(env, OGB_RANK=str(r))) for r in range(a.gpus)]
    rc = 0
    t0 = time.time()
    for p in procs:
        try:
            rc |= p.wait(timeout=max(1, 600 - (time.time() - t0)))
        except subprocess.TimeoutExpired:
            p.kill(); rc |= 1
    sys.exit(rc)


if __name__ == "__main__":
    worker() if os.environ.get("OGB_WORKER") else main()
