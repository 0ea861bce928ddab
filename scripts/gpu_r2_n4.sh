# round 2: contig-partitioned multi-GPU on 4 GPUs: group tests, `coverm --gpus 4`, bench --gpus 2 under torchrun
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_gpu_multi_n4.log 2>&1; echo "pytest multi rc=$?"; tail -15 gpurun_out/r2_gpu_multi_n4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.log; echo "bench n2 rc=$?"
tail -5 gpurun_out/r2_bench_n4.log
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n4.json'))
print('value',d['value'],'ms',d['ms_per_step'],'scaling',d['scaling'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'same',d['e2e']['output_identical_to_single_gpu_run'],'cold',d['e2e']['cold_cli'])
print(d['device_breakdown_ms_rank0']); print(d['e2e']['breakdown_last_step_rank0'])
P
