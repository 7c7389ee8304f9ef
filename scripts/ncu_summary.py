import csv, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,data=rows[0],rows[1],rows[2:]
keys=['gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor',
'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    for i,h in enumerate(hdr):
        if h==k: print(f"{k} [{units[i]}] = {[d[i] for d in data]}")
for i,h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        v=[float(d[i]) for d in data]
        if max(v)>0.05: print(h.replace('smsp__average_warps_issue_stalled_','stall ').replace('_per_issue_active.ratio',''), v)
