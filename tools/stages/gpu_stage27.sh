#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -c 5 -o gpurun_out/prof_gemm_final \
   env B2B_STEPS=1 python tools/decode_eager.py > gpurun_out/ncu_gemm_final.log 2>&1; echo "ncu exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_final.csv \
   env B2B_STEPS=1 python tools/decode_eager.py > gpurun_out/ncu_launches_final.log 2>&1; echo "launches exit $?"
