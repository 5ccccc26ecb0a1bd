#!/usr/bin/env python
"""Summaries of ncu outputs (launch list csv / raw page csv) for profiles/*.md."""
import csv, collections, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[h + 2:]:
        if len(r) <= vi: continue
        n = r[ki].split('(')[0][:70]
        agg.setdefault(n, []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    out = []
    for n, v in agg.items():
        out.append(f"{n:70s} n={len(v):3d} avg={sum(v)/len(v)/1000:8.1f} us share={sum(v)/tot*100:5.1f}%")
    return "\n".join(out)

def raw(path, keys=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    keys = keys or ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
                    'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
                    'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
    out = []
    for r in rows[2:]:
        out.append(r[idx['Kernel Name']].split('(')[0])
        for k in keys:
            if k in idx: out.append(f"    {k:70s} {r[idx[k]]} {rows[1][idx[k]]}")
    return "\n".join(out)

def traffic(path):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by bench.py's stage names -> profiles/dram_traffic.json"""
    import json
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    stage = {'preprocess_kernel': 'preprocess', 'duplicate_kernel': 'duplicate', 'tile_sort_small_kernel': 'tile_sort',
             'blend_forward_kernel': 'blend_forward', 'blend_backward_kernel': 'blend_backward',
             'gauss_backward_kernel': 'gauss_backward', 'tile_scan_kernel': 'scan'}
    agg = {}
    for r in rows[2:]:
        n = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('gsb::', '').split('<')[0]
        if n not in stage: continue
        b = sum(float(r[idx[k]].replace(',', '')) * unit[rows[1][idx[k]]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        agg.setdefault(stage[n], []).append(b)
    return json.dumps({k: sum(v) / len(v) for k, v in agg.items()}, indent=1)

def instructions(path):
    """smsp__inst_executed.sum (warp instructions) per launch, keyed like traffic() -> profiles/ncu_inst_executed.json"""
    import json
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    stage = {'preprocess_kernel': 'preprocess', 'duplicate_kernel': 'duplicate', 'tile_sort_small_kernel': 'tile_sort',
             'blend_forward_kernel': 'blend_forward', 'blend_backward_kernel': 'blend_backward',
             'gauss_backward_kernel': 'gauss_backward'}
    agg = {}
    for r in rows[2:]:
        n = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('gsb::', '').split('<')[0]
        if n in stage:
            agg.setdefault(stage[n], []).append(float(r[idx['smsp__inst_executed.sum']].replace(',', '')))
    return json.dumps({k: sum(v) / len(v) for k, v in agg.items()}, indent=1)

if __name__ == '__main__':
    cmd = sys.argv[1]
    print(launches(sys.argv[2]) if cmd == 'launches' else traffic(sys.argv[2]) if cmd == 'traffic' else instructions(sys.argv[2]) if cmd == 'instructions' else raw(sys.argv[2]))
