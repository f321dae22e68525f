mkdir -p gpurun_out
timeout 900 python tests/drivers/far_variants.py 0 13 22 41 > gpurun_out/r02c_far_variants.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02c_pytest_parity_v0.log 2>&1
BEMB200_FAR_VARIANT=13 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02c_pytest_parity_v13.log 2>&1
BEMB200_FAR_VARIANT=13 timeout 600 ncu --set full --clock-control none --import-source on -k regex:far_kernel -s 1 -c 1 -f -o gpurun_out/r02c_far_v13 python tests/drivers/far_only.py > gpurun_out/r02c_ncu_far_v13.log 2>&1
tail -12 gpurun_out/r02c_far_variants.log; tail -3 gpurun_out/r02c_pytest_parity_v0.log gpurun_out/r02c_pytest_parity_v13.log
