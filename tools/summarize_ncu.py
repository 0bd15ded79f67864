#!/usr/bin/env python3
"""Summarise ncu outputs into the small CSV/markdown files committed under profiles/.
  python tools/summarize_ncu.py launches <launches.csv>           -> per-kernel totals and shares
  python tools/summarize_ncu.py full <report.ncu-rep> <out.csv>   -> selected metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__inst_issued.avg.per_cycle_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.max', 'sm__cycles_elapsed.avg', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'device__attribute_multiprocessor_count',
        'sm__pipe_fp16_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        name = r[ki].split('(')[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.3f | %.1f |" % (k, v[0], v[1] / 1e3, v[1] / tot, v[1] / v[0] / 1e3))
    # the front kernel's share of ONE DDC step (what bench.py reports as roofline.kernel_share_of_step): the launch list holds
    # both workloads and the peak microbenchmarks, so the share is taken over the steps of the 1024-channel DDC workload -
    # consecutive prepare / front / ciccomp / hilb / rotate launches with the shorter front kernel
    seq, order = [], ("adc_prepare", "ddc_front_tc", "ddc_ciccomp", "ddc_hilb", "ddc_rotate")
    named = []
    for r in rows[1:]:
        try:
            named.append((r[ki].split('(')[0].replace('void ', ''), float(r[vi].replace(',', '')) / 1e3))
        except ValueError:
            pass
    for i in range(len(named) - 4):
        if all(named[i + j][0].startswith(order[j]) for j in range(5)):
            seq.append([named[i + j][1] for j in range(5)])
    if seq:
        short = min(s[1] for s in seq) * 1.5
        seq = [s for s in seq if s[1] < short]
        m = [sum(s[j] for s in seq) / len(seq) for j in range(5)]
        print("\nDDC workload, %d steps: prepare %.1f + front %.1f + ciccomp %.1f + hilb %.1f + rotate %.1f = %.1f us per step; "
              "front kernel share %.3f" % ((len(seq),) + tuple(m) + (sum(m), m[1] / sum(m))))


def full(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    with open(out, 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(rows) - 2)])
        for k in KEEP:
            if k in h:
                i = h.index(k)
                w.writerow([k, rows[1][i]] + [r[i] for r in rows[2:]])
    print(open(out).read())


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3])
