import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from metagenomics_b200 import Dataset, synth
from metagenomics_b200.dist import make_context
from metagenomics_b200._lib import check, lib
from metagenomics_b200.api import EDGE_DTYPE
ctx, rank, world, local = make_context()
L = lib()
cfg = synth.config(2, scale=1.0*world)
ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=50)
n = ds.getNumberOfUniqueReads(); words, woffs, lens = ds.packed()
def pinned(a):
    p = C.c_void_p(); check(L.ogb_alloc_host(C.byref(p), max(a.nbytes,1))); C.memmove(p, a.ctypes.data, a.nbytes); return p
pw, po, pl = pinned(words), pinned(woffs), pinned(lens)
pe = C.c_void_p(); check(L.ogb_alloc_host(C.byref(pe), 200<<20))
for it in range(4):
    torch.distributed.barrier() if world>1 else None
    t0=time.perf_counter(); check(L.ogb_reads_upload_packed(ctx._h, pw, po, pl, n)); t1=time.perf_counter()
    check(L.ogb_hash_build(ctx._h, 50)); t2=time.perf_counter()
    check(L.ogb_mark_contained(ctx._h)); check(L.ogb_build_graph(ctx._h, 0)); t3=time.perf_counter()
    ne=C.c_uint64(); check(L.ogb_graph_edge_count(ctx._h,0,C.byref(ne))); check(L.ogb_graph_edges(ctx._h,0,pe,ne.value)); t4=time.perf_counter()
    print(f"rank {rank} it {it}: upload {1e3*(t1-t0):.2f} hash {1e3*(t2-t1):.2f} build {1e3*(t3-t2):.2f} fetch {1e3*(t4-t3):.2f} ms  stats_total {ctx.stats()['ms_total']:.2f}", flush=True)
ctx.close()
