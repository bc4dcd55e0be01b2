"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path -- shard bounds, the unique-id
broadcast, max-over-ranks timing -- and a numpy emulation of the sharded algorithm (each rank scans
its own reads -> allgather of the adjacency -> each rank marks its own nodes -> allgather of the flags
-> twin merge) that must reproduce the oracle's post-reduction graph. The CUDA/NCCL path implements
exactly this decomposition (csrc/ogb_device.cu, ogb_build_graph)."""
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


def mark_own_nodes(adj, lo, hi):
    """SURVEY.md App. A.4 on the nodes (lo, hi] given the full pre-reduction adjacency {u: [(offset,dst,orient)] sorted}."""
    flags = {}
    for u in range(lo + 1, hi + 1):
        g = adj.get(u, [])
        state = {e[1]: 1 for e in g}
        for off, v, t1 in g:
            if state[v] != 1:
                continue
            for _, w, t2 in adj.get(v, []):
                if state.get(w) == 1 and compatible(t1, t2):
                    state[w] = 2
        flags[u] = [state[e[1]] == 2 for e in g]
    return flags


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
        for cfg in (datasets.small_configs()[0], datasets.tandem(mixed=True), datasets.palindromes()):
            orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"]).run_all(Oracle.BFS)
            n = orc.n
            lo, hi = ogd.shard_bounds(n, rank, world)
            bounds = [ogd.shard_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            pre = orc.edges(pre=True)                                  # (src,dst,offset,orient), canonical order
            mine = pre[(pre[:, 0] > lo) & (pre[:, 0] <= hi)]           # what this rank's scan (K3) emits
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)                     # C1
            full = np.concatenate(gathered)
            assert np.array_equal(full, pre)
            adj = {}
            for s, d, off, o in full.tolist():
                adj.setdefault(s, []).append((off, d, o))
            own_flags = mark_own_nodes(adj, lo, hi)                    # K5 on own nodes
            allflags = [None] * world
            dist.all_gather_object(allflags, own_flags)                # C2
            flags = {}
            for f in allflags:
                flags.update(f)
            keep = []                                                  # K6: survives iff neither side flagged it
            for u in range(lo + 1, hi + 1):
                for k, (off, w, o) in enumerate(adj.get(u, [])):
                    if flags[u][k]:
                        continue
                    tw = next(i for i, e in enumerate(adj[w]) if e[1] == u)
                    if not flags[w][tw]:
                        keep.append((u, w, off, o))
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
