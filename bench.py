#!/usr/bin/env python
"""bench.py -- overlap-graph build throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2] [--scale 1.0]

A *step* is one pass of the hot path over the workload with the packed reads already resident in
HBM: K1 hash build -> K2 containment (skipped for one read length, like the reference) -> K3 window
scan + verification -> K5 transitive marking -> K6 twin merge + compaction (+ the NCCL exchanges on
N > 1). `value` = unique reads / device time of a step (CUDA events on the library's stream, mean of
K steps, max over ranks), L2 flushed between steps. `e2e` is the same metric through the public C
ABI with HOST buffers: H2D of the packed reads from pinned memory + build + D2H of the final edge
list, every step. `roofline` is for the dominant kernel (K3); `cpu_baseline` times the reference's own
CPU implementation (oracle/_ref/ref_overlap, the unmodified reference) on a bounded sample.

--impl reference runs that CPU implementation as the measured arm (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "overlap_graph_build_reads_per_sec"
UNIT = "reads/s"
SECTOR = 32           # bytes per index bucket / DRAM sector
EDGE_BYTES = 8        # device edge record (offset<<48 | dst<<16 | orient<<8)
CPU_SAMPLE_SCALE = {1: 1.0, 2: 0.12, 3: 0.015, 4: 0.003, 5: 0.08}   # bounded sample: ~10-30 s of single-core reference


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md): a thread polling NVML every 2 ms
    (the timed region of this bench lasts tens of milliseconds -- `nvidia-smi -lms` would not produce a line in time)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, device):
        self.sm, self.mask, self.h, self.nv, self.run, self.mx = [], 0, None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _reasons(self):
        for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            f = getattr(self.nv, name, None)
            if f is not None:
                try:
                    return int(f(self.h))
                except Exception:
                    pass
        return 0

    def _poll(self):
        while self.run:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= self._reasons()
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self.run = True
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.run = False
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(self.sm)}


def workload(cfg_id, scale, n_gpus):
    from metagenomics_b200 import synth
    s = scale * n_gpus          # weak scaling: the genome and the read count grow with the GPU count
    cfg = synth.config(cfg_id, scale=s)
    name = cfg["name"] + (f" x{n_gpus} (weak scaling: {n_gpus}x genome and reads)" if n_gpus > 1 else "") + (f" [scale {scale}]" if scale != 1.0 else "")
    return cfg, name


def cpu_reference(cfg_id, steps, warmup, scale=None):
    """Times the reference's own CPU implementation (unmodified, via oracle/_ref/ref_overlap; falls back
    to the oracle port) on a bounded sample of the workload. Returns (reads_per_s, info dict)."""
    from metagenomics_b200 import synth
    from oracle_lib import Oracle, have_reference, run_reference
    sc = scale if scale is not None else CPU_SAMPLE_SCALE[cfg_id]
    cfg = synth.config(cfg_id, scale=sc)
    n_raw = len(cfg["offsets"]) - 1
    sample = f"{cfg['name']} at scale {sc} ({n_raw} raw reads, same coverage)"
    times, n_unique, e_final = [], 0, 0
    if have_reference():
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "sample.fa")
            synth.write_fasta(fa, cfg["bases"], cfg["offsets"])
            for i in range(warmup + steps):
                d, t, _ = run_reference([fa], cfg["min_overlap"], paired=cfg["paired"])
                if i >= warmup:
                    times.append(t["t_insert_s"] + t["t_build_s"])     # insertDataset + build to :210, mate-pair I/O excluded
                n_unique, e_final = d["n"], len(d["edges"])
        kind, cores = "reference", 1
    else:
        cores = os.cpu_count() or 1
        for i in range(warmup + steps):
            orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"])
            t0 = time.perf_counter()
            orc.run_all(Oracle.THREE_PHASE, threads=cores)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            n_unique, e_final = orc.n, len(orc.edges())
        kind = "port"
    t = float(np.mean(times))
    return n_unique / t, dict(kind=kind, cores=cores, sample=sample, seconds_per_step=t, n_unique=n_unique, edges_final=e_final,
                              edges_per_s=e_final / t)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    v, info = cpu_reference(args.config, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": {"workload": info["sample"]},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "edges_per_sec": info["edges_per_s"], "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="also check the result against the oracle (small scales)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import __graft_entry__ as g
    g.build(quiet=True)

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from metagenomics_b200 import Context, Dataset, nccl_unique_id
    from metagenomics_b200._lib import check, lib
    from metagenomics_b200.api import EDGE_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the overlap-graph build has no CPU path")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    ctx = Context(local_rank, rank, world, uid)
    L = lib()

    # ---- setup (untimed): synthetic reads -> host Dataset stage (filter, canonical strand, sort, dedupe)
    cfg, wname = workload(args.config, args.scale, world)
    t0 = time.perf_counter()
    ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
    t_dataset = time.perf_counter() - t0
    # the same stage with canonical strand / sort / dedupe on the GPU (second call: CUDA modules and pools warm), for the record
    t_dataset_dev = None
    if rank == 0 and world == 1:
        for _ in range(2):
            t0 = time.perf_counter()
            dsd = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"], device=ctx)
            t_dataset_dev = time.perf_counter() - t0
        assert dsd.getNumberOfUniqueReads() == ds.getNumberOfUniqueReads()
        del dsd
    n_unique = ds.getNumberOfUniqueReads()
    words, woffs, lens = ds.packed()
    m = cfg["min_overlap"]

    # pinned host staging for the end-to-end arm
    def pinned(arr):
        p = C.c_void_p()
        check(L.ogb_alloc_host(C.byref(p), max(arr.nbytes, 1)))
        C.memmove(p, arr.ctypes.data, arr.nbytes)
        return p
    p_words, p_offs, p_lens = pinned(words), pinned(woffs), pinned(lens)
    # a uniform-length, tightly packed input travels as words only (ogb_reads_upload_packed derives the rest)
    uniform = n_unique > 0 and int(lens.min()) == int(lens.max())
    h2d = words.nbytes + (0 if uniform else woffs.nbytes + lens.nbytes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def upload():
        check(L.ogb_reads_upload_packed(ctx._h, p_words, p_offs, p_lens, n_unique))

    def build():
        check(L.ogb_hash_build(ctx._h, m))
        check(L.ogb_mark_contained(ctx._h))
        check(L.ogb_build_graph(ctx._h, 0))

    upload()
    build()                                   # sizes every pool; later steps allocate nothing
    st = ctx.stats()
    n_final = st["edges_final"]
    p_edges = C.c_void_p()
    check(L.ogb_alloc_host(C.byref(p_edges), max(n_final, 1) * EDGE_DTYPE.itemsize))
    d2h = n_final * EDGE_DTYPE.itemsize
    FLUSH = 512 << 20

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            check(L.ogb_l2_flush(ctx._h, FLUSH))
            fn()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        per, wall0 = [], time.perf_counter()
        for _ in range(steps):
            check(L.ogb_l2_flush(ctx._h, FLUSH))      # untimed: the step starts with a cold L2
            barrier()
            ms = C.c_float()
            check(L.ogb_timer_begin(ctx._h))
            fn()
            check(L.ogb_timer_end(ctx._h, C.byref(ms)))
            per.append(ms.value)
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor(per, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks, per step
        return t.cpu().numpy(), wall, clocks

    # ---- value: device-resident inputs
    per_step, wall, clocks = timed(build, args.steps, args.warmup)
    st = ctx.stats()
    ms_step = float(per_step.mean())
    value = n_unique / (ms_step * 1e-3)

    # ---- e2e: host buffers in, host edge list out, through the C ABI
    def e2e_step():
        upload()
        build()
        n = C.c_uint64()
        check(L.ogb_graph_edge_count(ctx._h, 0, C.byref(n)))
        check(L.ogb_graph_edges(ctx._h, 0, p_edges, n.value))
    e2e_steps, _, _ = timed(e2e_step, max(2, args.steps // 2 + 1), 2)
    e2e_ms = float(e2e_steps.mean())

    if args.verify and rank == 0:
        from oracle_lib import Oracle, sort_tuples
        from metagenomics_b200 import edges_as_tuples
        got = np.ctypeslib.as_array(C.cast(p_edges, C.POINTER(C.c_uint8)), shape=(n_final * 12,)).view(EDGE_DTYPE)
        orc = Oracle(cfg["bases"], cfg["offsets"], m).run_all(Oracle.THREE_PHASE, threads=os.cpu_count() or 1)
        assert np.array_equal(sort_tuples(edges_as_tuples(got)), orc.edges()), "bench result differs from the oracle"

    # Roofline of the dominant kernel = the probe kernel (k_probe_uniform / k_probe), one launch per
    # chunk of query reads. Algorithmic bytes of one launch (DESIGN.md "Roofline", SURVEY.md 8(d) K3
    # terms): one 32-byte sector per probed window + the 12-byte candidate record per fingerprint match
    # + the query strand of every read of the chunk streamed once.
    read_bytes = int(np.mean([(int(l) + 63) // 64 * 16 for l in lens[:: max(1, len(lens) // 4096)]])) if n_unique else 32
    launches = max(1, st["probe_launches"])
    from metagenomics_b200.dist import shard_bounds
    s_lo, s_hi = shard_bounds(n_unique, rank, world)
    scan_bytes = (SECTOR * st["overlap_probes"] + 12 * st["candidates"] + read_bytes * (s_hi - s_lo)) / launches
    peak, peak_src = measured_peaks()

    if rank == 0:
        achieved = scan_bytes / (st["ms_probe_launch"] * 1e-3) / 1e9 if st["ms_probe_launch"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("workload") == f"config{args.config}@{args.scale}x{world}":
                traffic = tj.get("dram_bytes_per_launch")
        mark_bytes = 8.0 * (st["edges_pre_local"] + st["pivot_entries"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": wname, "min_overlap": m, "raw_reads": len(cfg["offsets"]) - 1, "unique_reads": n_unique,
                       "l2": "flushed between steps (512 MiB write, untimed)", "parallelism": f"query-read shards x{world}, replicated index",
                       "timing": "CUDA events on the library stream per step; max over ranks; mean of steps"},
            "edges_per_sec": st["edges_final"] / (ms_step * 1e-3), "edges_pre_per_sec": st["edges_pre"] / (ms_step * 1e-3),
            "edges_final": st["edges_final"], "edges_pre": st["edges_pre"],
            "e2e": {"value": n_unique / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "per_step_ms": [round(float(x), 3) for x in e2e_steps]},
            "per_step_ms": [round(float(x), 3) for x in per_step],
            "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "clocks": clocks,
            # Dominant single kernel of a step (ncu launch list profiles/launches_r1b.csv: 29 % of the GPU time): K5 k_mark, one launch per step.
            # Algorithmic bytes (SURVEY.md 8(d), K5 term): (E_pre + T) * B_e -- every node's own list plus the list of each active pivot, 8-byte edge words.
            "roofline": {"kernel": "k_mark (K5 markTransitiveEdges: warp per node on the unsorted slot regions, pivot order by warp min-reduction)",
                         "bound": "hbm", "achieved": mark_bytes / (st["ms_mark"] * 1e-3) / 1e9 if st["ms_mark"] > 0 else 0.0, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": (mark_bytes / (st["ms_mark"] * 1e-3) / 1e9 / peak) if st["ms_mark"] > 0 else 0.0,
                         "traffic": traffic.get("k_mark") if traffic else None, "algorithmic_bytes_per_launch": mark_bytes, "kernel_ms": st["ms_mark"],
                         "launches_per_step": 1, "share_of_step": st["ms_mark"] / ms_step if ms_step else None,
                         "note": "random 208-byte list fetches (~2.6 pivots per node) from GB-sized slot regions: 62-68 % of the issue slots busy, DRAM at ~20 % of peak -- issue/latency-bound, not bandwidth-bound (profiles/r1_notes.md, profiles/prof_r1b_reduce_summary.txt)"},
            # The K3 probe of one chunk (two kernels back to back), same definition as in the first half of the round.
            "roofline_k3_probe": {"kernel": "k_window_part (window hash + summary filter + scatter to partition queues) + k_probe_parts (64-byte bucket gather + fingerprint match), per chunk of query reads",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic.get("k3_probe") if traffic else None, "algorithmic_bytes_per_launch": scan_bytes, "kernel_ms": st["ms_probe_launch"],
                         "launches_per_step": launches, "share_of_step": st["ms_probe_launch"] * launches / ms_step if ms_step else None,
                         "note": "kernel_ms is measured in place while the previous chunk's verify runs on the second stream (alone: 0.23 ms per 256 k reads); the index partition being probed is L2-resident"},
            "phases_ms": {k: st[k] for k in ("ms_hash_build", "ms_contain", "ms_overlap", "ms_scan_kernel", "ms_exchange_pre", "ms_mark", "ms_reduce", "ms_total")},
            "stats": {k: st[k] for k in ("table_bytes", "overlap_probes", "probe_sectors", "candidates", "pivot_entries", "active_pivots",
                                         "max_degree", "overflow_reads", "n_contained", "nodes_final")},
            "setup_s": {"dataset_sort_dedupe": t_dataset, "dataset_sort_dedupe_device": t_dataset_dev}, "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline and world == 1:
            v, info = cpu_reference(args.config, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                                    "seconds": info["seconds_per_step"], "edges_per_sec": info["edges_per_s"]}
        print(json.dumps(line), flush=True)

    for p in (p_words, p_offs, p_lens, p_edges):
        L.ogb_free_host(p)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
