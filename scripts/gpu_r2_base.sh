# round 2 baseline: GPU tests + bench (N=1, config 2) + reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
cat /sys/fs/cgroup/cpu.max /sys/fs/cgroup/memory.max 2>/dev/null; nproc; free -g | head -2; df -h /tmp | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_a.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.log; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_a.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'cold',d['e2e']['cold_cli'],'cpu',d['cpu_baseline'],'parity',d['parity'])
print(d['device_breakdown_ms']); print(d['e2e']['breakdown_last_step'])
P
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.log; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_bench_ref_a.json
