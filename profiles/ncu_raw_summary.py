import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; data=rows[2:]
idx={h:i for i,h in enumerate(hdr)}
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__registers_per_thread','launch__grid_size','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.per_cycle_active','sm__inst_executed_pipe_lsu.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_red.sum','lts__t_sectors_srcunit_tex_op_atom.sum','l1tex__m_xbar2l1tex_read_sectors.sum','smsp__inst_executed_op_global_red.sum','smsp__inst_executed_op_global_st.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
for w in want:
    if w in idx:
        vals=[d[idx[w]] for d in data]
        if w=='Kernel Name': vals=[v[v.find('<'):v.find('>')+1] for v in vals]
        print('%-88s'%w, vals, units[idx[w]])
