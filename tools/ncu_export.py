import csv,sys,subprocess
# compact per-kernel ncu summary (selected raw metrics) -> CSV under profiles/
keep=('gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum')
out=csv.writer(open(sys.argv[1],'w'))
out.writerow(['report','kernel','metric','unit','value'])
for rep in sys.argv[2:]:
    txt=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(txt.splitlines()))
    hdr,units,vals=rows[0],rows[1],rows[2]
    kn=vals[hdr.index('Kernel Name')]
    for i,h in enumerate(hdr):
        if h in keep or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            out.writerow([rep.split('/')[-1],kn[:60],h,units[i],vals[i]])
