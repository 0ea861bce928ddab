# round 2: K2 warp-cooperative sparse path A/B, coverm filter on the device, filter.rs goldens, e2e outlier diagnosis
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filter or golden or synthetic_bams" > gpurun_out/r2_gpu_tests_k2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests_k2.log
run() { name=$1; shift
  timeout 600 python bench.py --steps 8 --warmup 3 --skip-cold-cli --skip-cpu-baseline "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.log; echo "bench $name rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_$name.json'))
print('$name value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'])
print(d['device_breakdown_ms_rank0']); print([round(x,3) for x in d['e2e']['step_walls_s']])
P
  grep "e2e per step" gpurun_out/r2_bench_$name.log
}
run k2coop
run k2nocoop --lib variants/k2nocoop.so
