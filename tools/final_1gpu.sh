#!/bin/bash
# Round-end evidence on one B200: bench line, CPU arm, ncu --set full of the C2 step, launch list.
set -x
t0=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$? after $(( $(date +%s) - t0 )) s"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
timeout 120 python tools/prof_step.py C2 1.0 4 > gpurun_out/plain_e.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sp_wtile|sp_split|sp_group|bin_mean|reads_to_global|sp_chunk" -s 12 -c 6 -f -o gpurun_out/prof_r02e python tools/prof_step.py C2 1.0 4 > gpurun_out/ncu_e.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_e_launches_C2.csv python tools/prof_step.py C2 1.0 3 > gpurun_out/ncu_e2.log 2>&1
tail -n 2 gpurun_out/plain_e.log gpurun_out/ncu_e.log; tail -c 400 gpurun_out/bench_final.err
