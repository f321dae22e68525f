#!/bin/bash
# round-2 profile pass on one B200: ncu --set full of the new kernels, launch list of the bench command
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:zgemm_streamk -c 1 -o gpurun_out/r02l_streamk python tests/drivers/block_matvec_only.py 5 32 2 > gpurun_out/r02l_ncu1.log 2>&1
PROBE_SIZES=64 PROBE_KINDS=contiguous timeout 200 $NCU -k regex:schwarz_apply -c 1 -o gpurun_out/r02l_schwarz_apply python tests/drivers/precond_probe.py 5 2 > gpurun_out/r02l_ncu2.log 2>&1
PROBE_SIZES=256 PROBE_KINDS=contiguous timeout 200 $NCU -k "regex:lu_nopivot|invert_kernel" -c 2 -o gpurun_out/r02l_schwarz_setup python tests/drivers/precond_probe.py 5 2 > gpurun_out/r02l_ncu3.log 2>&1
BENCH_NO_CONFIG5=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02l_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_ncu5.log 2>&1
ls -la gpurun_out/r02l_* | head -20
