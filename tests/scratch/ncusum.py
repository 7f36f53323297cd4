import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units=rows[0],rows[1]
keys=['gpu__time_duration.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','launch__block_size']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('==',d.get('Kernel Name','')[:90])
    for k in keys:
        if k in d: print(f'  {k} = {d[k]} {units[hdr.index(k)]}')
    for h in hdr:
        if 'issue_stalled' in h and 'per_issue_active' in h:
            v=float(d[h].replace(',','')) if d[h] else 0
            if v>0.15: print(f'  stall {h.split("issue_stalled_")[1].split("_per_issue")[0]} = {v:.2f}')
