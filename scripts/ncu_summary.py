"""Prints the metrics that matter for the traversal kernels from an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__sass_average_branch_targets_threads_uniform.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__inst_executed.avg.per_cycle_active',
        'sm__sass_inst_executed_op_local_ld.sum', 'sm__sass_inst_executed_op_local_st.sum',
        'sm__sass_inst_executed_op_shared_ld.sum', 'sm__sass_inst_executed_op_shared_st.sum',
        'sm__sass_inst_executed_op_global_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
idx = {h: i for i, h in enumerate(hdr)}
for w in want:
    if w in idx:
        print(w.ljust(72), [r[idx[w]] for r in rows[2:]], units[idx[w]])
print("-- stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active), >= 0.1")
for h in hdr:
    if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
        vals = [r[idx[h]] for r in rows[2:]]
        try:
            if max(float(v.replace(',', '')) for v in vals) >= 0.1:
                print(h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')].ljust(30), vals)
        except ValueError:
            pass
print("-- L1 by memory space")
for w in ['l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct',
          'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
          'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']:
    if w in idx:
        print(w.ljust(72), [r[idx[w]] for r in rows[2:]], units[idx[w]])
