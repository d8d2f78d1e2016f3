M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct,smsp__warps_eligible.avg.per_cycle_active
for opt in "extend_defer=0" "extend_defer=1"; do
  PTB_OPTIONS=$opt,extend_rays_per_lane=0 ncu --metrics $M --clock-control none -k regex:extend_lanes -s 4 -c 4 --csv --log-file gpurun_out/ab_$opt.csv python scripts/profile_extend.py > /dev/null 2>&1
done
