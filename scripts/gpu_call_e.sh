mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02e_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_solver_callers.py "tests/test_gpu_parity.py::test_far_kernel_math" -x -q -m gpu > gpurun_out/r02e_pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err; echo "bench2 rc=$?"
tail -n 4 gpurun_out/r02e_pytest_multi_2gpu.log; tail -n 5 gpurun_out/r02e_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_2gpu.json'))
print(d['value'], d['e2e']['value'], d['run']['solver'], d['run']['exchange'])
print(json.dumps(d.get('sharded_parity')))
print(json.dumps(d.get('config3'))[:3000])
PY
