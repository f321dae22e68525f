mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02d_bench_1gpu_default.json 2> gpurun_out/r02d_bench_1gpu_default.err; echo "bench rc=$?"
BEMB200_GMRES_FUSED=1 BENCH_NO_CONFIG5=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench_1gpu_fused_pipelined.json 2> gpurun_out/r02d_bench_1gpu_fused_pipelined.err; echo "bench fused pipelined rc=$?"
BEMB200_GMRES_FUSED=1 BENCH_NO_CONFIG5=1 timeout 600 python bench.py --no-cpu-baseline --schedule sequential > gpurun_out/r02d_bench_1gpu_fused_sequential.json 2> gpurun_out/r02d_bench_1gpu_fused_sequential.err; echo "bench fused seq rc=$?"
BENCH_NO_CONFIG5=1 timeout 600 python bench.py --no-cpu-baseline --schedule sequential > gpurun_out/r02d_bench_1gpu_sequential.json 2> gpurun_out/r02d_bench_1gpu_sequential.err; echo "bench seq rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:far_kernel -s 1 -c 1 -f -o gpurun_out/r02d_far python tests/drivers/far_only.py > gpurun_out/r02d_ncu_far.log 2>&1
tail -n 3 gpurun_out/r02d_pytest_gpu.log
for f in gpurun_out/r02d_bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', d['value'], d['e2e']['value'], d['run']['solver'], d['roofline']['avg_launch_ms'], d['roofline_assembly'].get('avg_launch_ms'), d.get('config5',{}).get('s_per_batch'))
"; done
