cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "$@"; do
  timeout 200 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --e2e-steps 1 --lib variants/$v/libcoverm_b200.so > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.log || { echo "$v FAILED/timeout"; tail -3 gpurun_out/bench_$v.log; continue; }
  python -c "import json; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', d['ms_per_step'], round(d['roofline']['frac'],4), d['device_breakdown_ms'])"
done
