mkdir -p gpurun_out
timeout 900 python tests/drivers/far_variants.py > gpurun_out/r02b_far_variants.log 2>&1
timeout 300 python tests/drivers/run_config5.py > gpurun_out/r02b_config5.log 2>&1
tail -30 gpurun_out/r02b_far_variants.log
