import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]; v=rows[-1]
d=dict(zip(h,v))
items=[(k,val) for k,val in d.items() if 'issue_stalled' in k and 'per_issue_active.ratio' in k]
print("stalls:", ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_','').replace('_per_issue_active.ratio','')}={float(val):.2f}" for k,val in sorted(items,key=lambda kv:-float(kv[1] or 0))[:8]))
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','lts__t_sectors.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed']:
    print(k, d.get(k))
