#!/bin/bash
# One GPU call: GPU tests, bench line, ncu launch list of the bench command, ncu --set full of the two kernels of the
# batch-256 update, in-kernel timelines (measurement script; outputs under gpurun_out/)
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
(time python bench.py) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"her_sample_kernel|ddpg_stream_kernel|rows_dw_kernel|tc_chain_kernel|tc_gemm_kernel|actions_stream_kernel|tc_chain_presplit|tc_chain_rowsum|prep_kernel|adam_kernel" -c 500 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ITERS=30 ncu --set full --clock-control none --import-source on -k regex:"ddpg_stream_kernel|rows_dw_kernel" -s 20 -c 4 -o gpurun_out/rows_prof python scratch/ddpg_bench.py > gpurun_out/rows_prof.log 2>&1
CUR_ROWS_TIMELINE=1 GRAPH=0 ITERS=100 python scratch/ddpg_bench.py > gpurun_out/rows_timeline_eager.log 2>&1
CUR_ROWS_TIMELINE=1 python scratch/ddpg_bench.py > gpurun_out/rows_timeline_graph.log 2>&1
python scratch/actions_bench.py > gpurun_out/actions.log 2>&1
ls -la gpurun_out | head -30
