#!/bin/bash
# per-kernel durations of the C1 solve (CGNR, 1024x4096 ComplexF32): ncu launch list, launch-by-launch path
RLS_SOLVE_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/r02_launches_c1.csv python tools/run_configs.py c1 > gpurun_out/c1_under_ncu.log 2>&1
