"""Summarise an .ncu-rep: per-kernel headline metrics (raw page) and, optionally, the hottest SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_fp32.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
units = rows[1]
for r in rows[2:]:
    print("-" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("  %-85s %s %s" % (w, r[i][:60], units[i]))
if kern:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    data = [(r[ia], int(r[ie]), int(r[isamp])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    n = len(data) // 2 if len(data) > 1 and data[0][0] == data[len(data) // 2][0] else len(data)
    data = data[:n]
    tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
    print("total warp instr", tot, "samples", ts)
    top = sorted(range(n), key=lambda i: -data[i][2])[:40]
    for i in sorted(top):
        print("%5d exec %12d  samp %5.1f%%  %s" % (i, data[i][1], 100.0 * data[i][2] / max(ts, 1), data[i][0][:100]))
