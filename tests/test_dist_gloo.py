"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path -- shard bounds, the unique-id
broadcast, max-over-ranks timing -- and a Python model of the sharded algorithm exactly as the CUDA/NCCL
path runs it (csrc/ogb_device.cu ogb_build_graph, csrc/ogb_kernels.cuh k_rows_finish / k_mark_fast / k_mark_any /
k_keep / k_emit): every rank keeps its own lists UNSORTED (discovery order), packs them to (dst, strand) entries
-- the adjacency rows -- that reach every rank (C1), walks the pivots of its own nodes by repeated minimum of
(offset, dst, orient, slot) among the in-play entries, records where the scan of a pivot met the twin entry, publishes
one ELIM bit per row entry (C2), keeps an edge iff the bit at the recorded twin position is clear (K6),
and the sorted survivors of all shards are allgathered (C3). The result must be the oracle's
post-reduction graph; the model also checks the claim K6 relies on: every edge a node keeps was one of
its pivots, so its twin position is always known."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compatible(t1, t2):
    return (t1 & 1) == ((t2 >> 1) & 1)


def mark_own_nodes(own, packed, lo, hi):
    """k_mark on the nodes (lo, hi]: own[u] = [(offset, dst, orient)] in discovery order, packed[v] = [dst<<1 | strand].
    Returns per node the ELIM verdicts and the twin positions (1-based, 0 = unknown), both in slot order."""
    elim, twin = {}, {}
    for u in range(lo + 1, hi + 1):
        g = own.get(u, [])
        state = {e[1]: 1 for e in g}                                   # all neighbours INPLAY (:577-578)
        tw = [0] * len(g)
        cur = None
        while True:
            cand = [((e[0], e[1], e[2]), k) for k, e in enumerate(g) if state[e[1]] == 1 and (cur is None or ((e[0], e[1], e[2]), k) > cur)]
            if not cand:
                break
            cur = min(cand)                                            # next pivot: smallest key above the current one still in play
            k = cur[1]
            _, v, t1 = g[k]
            for pos, f in enumerate(packed.get(v, [])):
                x, strand = f >> 1, f & 1
                if x == u:
                    tw[k] = pos + 1
                if (t1 & 1) == strand and state.get(x) == 1:
                    state[x] = 2                                       # :588-596
        elim[u] = [state[e[1]] == 2 for e in g]
        twin[u] = tw
    return elim, twin


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from metagenomics_b200 import dist as ogd
    import datasets
    from oracle_lib import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert ogd.env_rank() == (rank, world, rank)
        uid = ogd.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
        assert uid == bytes(range(128))
        assert ogd.max_over_ranks([1.0 + rank, 5.0 - rank]) == [float(world), 5.0]
        for cfg in (datasets.small_configs()[0], datasets.tandem(mixed=True), datasets.palindromes(), datasets.repeats(), datasets.even_h()):
            orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
            n = orc.n
            lo, hi = ogd.shard_bounds(n, rank, world)
            bounds = [ogd.shard_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            pre = orc.edges(pre=True)                                  # (src,dst,offset,orient), canonical order
            mine = pre[(pre[:, 0] > lo) & (pre[:, 0] <= hi)]           # what this rank's scan (K3) emits ...
            rng = np.random.default_rng(1000 + rank)
            own = {}
            for s_, d, off, o in mine[rng.permutation(len(mine))].tolist():   # ... in discovery order, i.e. unsorted
                own.setdefault(s_, []).append((off, d, o))
            mypacked = {u: [(d << 1) | ((o >> 1) & 1) for _, d, o in g] for u, g in own.items()}   # k_rows_finish
            gathered = [None] * world
            dist.all_gather_object(gathered, mypacked)                 # C1: packed lists (+ node records)
            packed = {}
            for part in gathered:
                packed.update(part)
            assert sum(len(v) for v in packed.values()) == len(pre)
            own_elim, own_twin = mark_own_nodes(own, packed, lo, hi)   # K5 on own nodes
            allbits = [None] * world
            dist.all_gather_object(allbits, own_elim)                  # C2: one ELIM bit per packed entry
            bits = {}
            for f in allbits:
                bits.update(f)
            keep = []                                                  # K6: survives iff neither side flagged it
            for u in range(lo + 1, hi + 1):
                surv = []
                for k, (off, w, o) in enumerate(own.get(u, [])):
                    if own_elim[u][k]:
                        continue
                    tw = own_twin[u][k]
                    assert tw > 0, "an edge its node keeps was a pivot of that node: the twin position is known"
                    assert packed[w][tw - 1] >> 1 == u
                    if not bits[w][tw - 1]:
                        surv.append(((off, w, o), k))
                keep += [(u, w, off, o) for (off, w, o), _ in sorted(surv)]   # k_emit: key order, slot as tie-break
            fin = [None] * world
            dist.all_gather_object(fin, keep)                          # C3
            got = np.array([e for part in fin for e in part], dtype=np.uint32).reshape(-1, 4)
            assert np.array_equal(got, orc.edges()), cfg["name"]
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
