mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_gpus.txt
timeout 900 python -m pytest "tests/test_gpu_solver_callers.py::test_single_process_group_two_ranks_on_one_gpu" -x -q -m gpu > gpurun_out/r02f_pytest_group_4gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_group_4gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 > gpurun_out/r02f_bench_4gpu.json 2> gpurun_out/r02f_bench_4gpu.err; echo "bench4 rc=$?"
tail -n 3 gpurun_out/r02f_pytest_group_4gpu.log; tail -n 3 gpurun_out/r02f_bench_4gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench_4gpu.json'))
print(d['value'], d['e2e']['value'], d['e2e'].get('sequential_value'), d['run']['solver'], d['run']['exchange'], d['roofline']['avg_launch_ms'], d['breakdown_ms_per_step'])
print(json.dumps(d.get('sharded_parity')))
print(json.dumps(d.get('config3'))[:3000])
PY
