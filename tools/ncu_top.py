"""Development helper: headline metrics + top stall lines (SASS with source line) of one kernel in an .ncu-rep."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:60])
    for w in want:
        if w in h: print("  ", w, r[h.index(w)], rows[1][h.index(w)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si = h.index('Source'); wi = h.index('Warp Stall Sampling (All Samples)'); ii = h.index('Instructions Executed'); ti = h.index('Avg. Threads Executed')
data = []
for r in rows[hi + 1:]:
    try: data.append((int(r[wi] or 0), int(r[ii] or 0), r[si], r[ti]))
    except Exception: pass
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print("samples", tot, "warp instr", toti)
for i in sorted(range(len(data)), key=lambda i: -data[i][0])[:topn]:
    d = data[i]; print(f"{i:5d} {d[0]*100/tot:5.1f}% inst={d[1]*100/toti:4.1f}% thr={d[3]:>5s} {d[2][:90]}")
