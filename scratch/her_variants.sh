#!/bin/bash
# HER kernel: time (Arm4, Arm8) + DRAM bytes / L1 / L2 sector counts of one launch (profiles/README.md, "HER kernel, round 2");
# CFGS="BUFSIZE=50000" keeps the buffers L2-resident (the SM-side floor of the kernel)
M=dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,smsp__inst_executed.sum
for cfg in ${CFGS:-"BUFSIZE=1000000" "BUFSIZE=1000000 NMOD=8"}; do
echo "== $cfg"
env $cfg python scratch/her_bench.py
env $cfg ITERS=3 ncu --metrics $M --clock-control none -k regex:her_sample -c 1 -s 5 python scratch/her_bench.py 2>&1 | grep -E "dram__|lts__|l1tex__|smsp__"
done
