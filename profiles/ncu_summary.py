#!/usr/bin/env python
"""Summarises an .ncu-rep: per-kernel headline metrics (raw page) and, with --source KERNEL, the
instruction / stall-sample share per CUDA source line. Usage:
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source k_scan] [--top 40]"""
import argparse, collections, csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def ncu(args):
    return subprocess.run(['ncu'] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def raw(rep):
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('-----', r[idx['Kernel Name']][:70])
        for w in WANT:
            if w in idx:
                print(f"  {w:82s} {r[idx[w]]} {units[idx[w]]}")


def source(rep, kernel, top, path):
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel, '--print-source', 'cuda,sass']))))
    secs = [i for i, r in enumerate(rows) if r and r[0] == 'File Path'] + [len(rows)]
    src = open(path).read().split('\n')
    for a, b in zip(secs[:-1], secs[1:]):
        if not rows[a][1].endswith(path.split('/')[-1]):
            continue
        hdr = rows[a + 2]
        ci, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
        agg = collections.OrderedDict(); text = {}
        for r in rows[a + 3:b]:
            if r[0] == '':
                continue
            try:
                agg[int(r[0])] = (int(r[ci]), int(r[si])); text[int(r[0])] = r[1]
            except ValueError:
                pass
        tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
        if tot < 1000:
            continue
        print(rows[a + 1][1][:60], 'instructions', tot, 'samples', tots)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / max(tots, 1) * 100:5.1f}% smp  L{k:>4} {text.get(k, '').strip()[:110]}")


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('rep'); ap.add_argument('--source'); ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--file', default='metagenomics_b200/csrc/ogb_kernels.cuh')
    a = ap.parse_args()
    if a.source:
        source(a.rep, a.source, a.top, a.file)
    else:
        raw(a.rep)
