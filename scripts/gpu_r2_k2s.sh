# round 2: K2 with a static chunk schedule and metadata requested an iteration ahead (variants/k2static.so) vs the default build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_$1.json'))
print('$1 value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'parity',d['parity'])
print(d['device_breakdown_ms_rank0'])
P
}
for v in main k2static main k2static; do
  L=""; [ $v != main ] && L="--lib variants/$v.so"
  timeout 600 python bench.py --steps 20 --warmup 3 --e2e-steps 1 --skip-cold-cli --skip-cpu-baseline $L > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.log; echo "bench $v rc=$?"; show $v
done
timeout 600 python bench.py --config ns --steps 5 --warmup 3 --e2e-steps 1 --skip-cold-cli --skip-cpu-baseline --lib variants/k2static.so > gpurun_out/r2_bench_k2static_ns.json 2> gpurun_out/r2_bench_k2static_ns.log; echo "bench ns k2static rc=$?"; show k2static_ns
cp coverm_b200/libcoverm_b200.so /tmp/main.so; cp variants/k2static.so coverm_b200/libcoverm_b200.so
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or synthetic_bams or fixtures or arena or retries" > gpurun_out/r2_gpu_parity_k2static.log 2>&1; echo "pytest k2static rc=$?"; tail -3 gpurun_out/r2_gpu_parity_k2static.log
cp /tmp/main.so coverm_b200/libcoverm_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inflate" > gpurun_out/r2_gpu_inflate_after_sync.log 2>&1; echo "pytest inflate rc=$?"; tail -2 gpurun_out/r2_gpu_inflate_after_sync.log
