#!/usr/bin/env python
"""one block per KERNEL of an .ncu-rep (ncu --set full): launches, total device time, and the counters of its longest launch
   python scripts/ncu_by_kernel.py rep.ncu-rep > profiles/<name>.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum']
ik, it = hdr.index('Kernel Name'), hdr.index('gpu__time_duration.sum')
def tms(r):
    try:
        v = float(r[it].replace(',', ''))
    except ValueError:
        return 0.0
    u = units[it]
    return v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u in ('us', 'usecond') else (v * 1e3 if u in ('s', 'second') else v))
by = {}
for r in rows[2:]:
    if len(r) <= it:
        continue
    by.setdefault(r[ik], []).append(r)
tot_all = sum(tms(r) for rs in by.values() for r in rs)
print("kernels: %d, launches: %d, total device time %.3f ms (ncu --set full, cold caches, serialised: compare SHARES)" % (len(by), sum(len(v) for v in by.values()), tot_all))
for name, rs in sorted(by.items(), key=lambda kv: -sum(tms(r) for r in kv[1])):
    tot = sum(tms(r) for r in rs)
    big = max(rs, key=tms)
    print('---')
    print('  %-72s %s' % ('Kernel Name', name[:140]))
    print('  %-72s %d launches, %.4f ms in total (%.1f %% of the capture); longest launch below' % ('launches', len(rs), tot, 100.0 * tot / max(tot_all, 1e-12)))
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print('  %-72s %s %s' % (w, big[i][:60], units[i]))
