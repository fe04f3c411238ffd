import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr,units,vals=rows[0],rows[1],rows[2]
keep=('gpu__time_duration','dram__bytes','dram__throughput','gpu__dram_throughput','pipe_fp64','sm__inst_executed','smsp__inst_executed.sum','issue_stalled','launch__','lts__t_sector_hit','l1tex__data_pipe_lsu_wavefronts_mem_shared','sass__inst_executed_shared','sm__throughput','sm__warps_active','smsp__issue_active','sm__cycles_elapsed.avg','smsp__thread_inst_executed','sm__sass_thread_inst_executed_op_d','Kernel Name','bank_conflict')
with open(sys.argv[2],'w') as f:
    w=csv.writer(f); w.writerow(['metric','unit','value'])
    for h,u,v in zip(hdr,units,vals):
        if any(k in h for k in keep): w.writerow([h,u,v])
