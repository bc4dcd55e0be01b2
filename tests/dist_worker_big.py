"""Worker of tests/test_multi_gpu.py (large cases): the ranks build one BASELINE.json configuration at a given scale through
the multi-rank C ABI -- rank 0 runs the Dataset stage, the packed reads reach the other hosts through /dev/shm, every rank
uploads its shard (replicated over NVLink) -- and compare counters and the order-independent checksum of the replicated final
edge list with the oracle's golden (tests/golden/full_size.json).

    torchrun ... tests/dist_worker_big.py <config> <scale>"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from metagenomics_b200 import Dataset, synth  # noqa: E402
from metagenomics_b200._lib import check, lib  # noqa: E402
from metagenomics_b200.dist import make_context, shared_packed_reads, upload_shared  # noqa: E402


def main():
    k, scale = int(sys.argv[1]), float(sys.argv[2])
    gold = [g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "full_size.json"))) if g["config"] == k and abs(g["scale"] - scale) < 1e-9]
    assert gold, f"no golden for config {k} @ {scale}"
    gold = gold[0]
    ctx, rank, world, local = make_context()
    torch.cuda.set_device(local)
    t0 = time.time()
    m = {1: 40, 2: 50, 3: 50, 4: 60, 5: 50}[k]

    def make():
        cfg = synth.config(k, scale=scale)
        return Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=m)
    words, woffs, lens, meta = shared_packed_reads(make, rank, world, tag="ogb_test")
    t1 = time.time()
    assert meta["n"] == gold["n_unique"], (meta["n"], gold["n_unique"])
    upload_shared(ctx, words, woffs, lens, meta, rank, world)
    L = lib()
    for rep in range(2):                                     # second build: pools sized, retry paths quiet
        check(L.ogb_hash_build(ctx._h, m))
        check(L.ogb_mark_contained(ctx._h))
        check(L.ogb_build_graph(ctx._h, 0))
    st = ctx.stats()
    x, s = C.c_uint64(), C.c_uint64()
    check(L.ogb_graph_checksum(ctx._h, 0, C.byref(x), C.byref(s)))
    tot = torch.tensor([st["pivot_entries"], st["edges_pre_local"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    T_all, E_all = (int(v) for v in tot.cpu())
    got = dict(E_pre=st["edges_pre"], E_final=st["edges_final"], nodes=st["nodes_final"], contained=st["n_contained"], T=T_all, checksum=[x.value, s.value])
    bad = {k_: (v, gold[k_]) for k_, v in got.items() if gold[k_] != v}
    assert not bad, (rank, bad)
    assert E_all == gold["E_pre"]
    print(f"rank {rank}/{world}: config {k} @ {scale} ({meta['n']} reads, {st['edges_pre']} -> {st['edges_final']} edges) identical to the golden; "
          f"setup {t1 - t0:.1f} s, step {st['ms_total']:.2f} ms [hash {st['ms_hash_build']:.2f} overlap {st['ms_overlap']:.2f} rows+C1 {st['ms_exchange_pre']:.2f} "
          f"mark {st['ms_mark']:.2f} reduce {st['ms_reduce']:.2f}]", flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
