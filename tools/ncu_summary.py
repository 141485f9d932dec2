"""Summarise ncu outputs from gpurun_out/ into profiles/ (text, committed).
usage: python tools/ncu_summary.py <launches.csv> <out.txt>            (launch list -> shares)
       python tools/ncu_summary.py --rep <file.ncu-rep> <out.txt>      (full capture -> key metrics)"""
import csv, subprocess, sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki]].append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    with open(out, 'w') as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)\n")
        f.write("# source: %s ; total %.3f ms over %d launches\n" % (path, tot / 1e6, sum(len(v) for v in agg.values())))
        f.write("%-110s %5s %12s %12s %7s\n" % ('kernel', 'n', 'mean_us', 'total_us', 'share'))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-110s %5d %12.1f %12.1f %7.3f\n" % (k[:110], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, sum(v) / tot))


def rep(path, out):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source: %s\n" % path)
        for r in rows[2:]:
            f.write("\n== %s  grid=%s block=%s\n" % (r[hdr.index('Kernel Name')], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("%-80s %-14s %s\n" % (k, units[i], r[i]))


if __name__ == '__main__':
    if sys.argv[1] == '--rep':
        rep(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[1], sys.argv[2])
