# usage: bash tools/_ncu_quick.sh name ...   (A/B builds under build/variants) -> gpurun_out/ncuq_<name>.csv
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tex.sum,sm__cycles_elapsed.avg,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
for n in "$@"; do
  DRR_B200_LIB=$PWD/build/variants/libdrr_$n.so ncu --metrics $M --clock-control none -k regex:march_warp -c 1 --csv --log-file gpurun_out/ncuq_$n.csv python tools/prof_one.py hybrid 1 4 1 > gpurun_out/ncuq_$n.log 2>&1
  python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncuq_$n.csv")) if len(r)>10]
h=rows[0]; i=h.index("Metric Name"); v=h.index("Metric Value")
print("== $n")
for r in rows[1:]: print("  %-90s %s"%(r[i],r[v]))
P
done
