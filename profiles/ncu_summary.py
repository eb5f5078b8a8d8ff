#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: one block per kernel launch with the metrics that matter
for an HBM-bound integer kernel. Usage: python profiles/ncu_summary.py raw.csv
       python profiles/ncu_summary.py --traffic KEY raw.csv   (KEY = WxH_nN_bB: adds the chain's measured DRAM bytes
       per launch to profiles/chain_traffic.json, which bench.py reads for roofline.traffic)"""
import csv
import json
import os
import sys


def traffic(key, path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        name = r[col['Kernel Name']].split('(')[0]
        rd = float(r[col['dram__bytes_read.sum']].replace(',', ''))
        wr = float(r[col['dram__bytes_write.sum']].replace(',', ''))
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        rd *= scale[rows[1][col['dram__bytes_read.sum']]]
        wr *= scale[rows[1][col['dram__bytes_write.sum']]]
        du = float(r[col['gpu__time_duration.sum']].replace(',', ''))
        du *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6, 'msecond': 1.0}[rows[1][col['gpu__time_duration.sum']]]
        per.setdefault(name, []).append((rd, wr, du))
    kern = {k: {'launches_captured': len(v), 'dram_read_bytes': sum(x[0] for x in v) / len(v),
                'dram_write_bytes': sum(x[1] for x in v) / len(v), 'ms_under_ncu': sum(x[2] for x in v) / len(v)}
            for k, v in per.items()}
    tot = sum(k['dram_read_bytes'] + k['dram_write_bytes'] for k in kern.values())
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'chain_traffic.json')
    tab = json.load(open(out)) if os.path.exists(out) else {}
    tab[key] = {'bytes_per_batch': tot, 'launches_per_batch': len(kern), 'bytes_per_launch': tot / len(kern),
                'kernels': kern, 'source': os.path.basename(path)}
    json.dump(tab, open(out, 'w'), indent=1, sort_keys=True)
    print(key, json.dumps(tab[key], indent=1))


if len(sys.argv) > 1 and sys.argv[1] == '--traffic':
    traffic(sys.argv[2], sys.argv[3])
    sys.exit(0)

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct',
        'smsp__warp_issue_stalled_sleeping_per_warp_active.pct', 'smsp__warp_issue_stalled_drain_per_warp_active.pct',
        'smsp__warp_issue_stalled_imc_miss_per_warp_active.pct', 'smsp__warp_issue_stalled_selected_per_warp_active.pct']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('----')
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"{w} = {r[i]} {units[i]}")
