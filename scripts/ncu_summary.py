"""Print the metrics that matter from an `ncu --page raw --csv` dump (one row per launch)."""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__inst_executed.sum', 'sm__inst_executed.sum.per_cycle_active', 'smsp__cycles_active.avg',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')][:90])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); print("  %-72s %-14s %s" % (k, units[i], r[i]))
    stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warp') and h.endswith('.ratio') and r[i] not in ('', 'n/a')]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print("  stall %-66s %.3f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''), v))
