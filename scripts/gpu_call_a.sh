mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02a_bench_1gpu_default.json 2> gpurun_out/r02a_bench_1gpu_default.err; echo "bench rc=$?"
timeout 300 python tests/drivers/far_only.py > gpurun_out/r02a_far_only.log 2>&1
timeout 300 python tests/drivers/far_sizes.py > gpurun_out/r02a_far_sizes.log 2>&1
timeout 300 python tests/drivers/fused_probe.py --big > gpurun_out/r02a_fused_probe_big.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --schedule sequential > gpurun_out/r02a_bench_seq.json 2> gpurun_out/r02a_bench_seq.err && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02a_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --schedule sequential > gpurun_out/r02a_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:far_kernel -s 1 -c 1 -f -o gpurun_out/r02a_far python tests/drivers/far_only.py > gpurun_out/r02a_ncu_far.log 2>&1
timeout 300 python tests/drivers/matvec_only.py > gpurun_out/r02a_matvec_only.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemv_kernel -s 2 -c 1 -f -o gpurun_out/r02a_zgemv python tests/drivers/matvec_only.py > gpurun_out/r02a_ncu_zgemv.log 2>&1
ls -la gpurun_out
