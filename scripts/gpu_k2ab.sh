cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/h2d_bw.py > gpurun_out/h2d_bw.log 2>&1; cat gpurun_out/h2d_bw.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "synthetic_bams or reuse or fixtures" > gpurun_out/gpu_tests_span32.log 2>&1
tail -3 gpurun_out/gpu_tests_span32.log
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_span32.json 2> gpurun_out/bench_span32.log
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --lib variants/span16/libcoverm_b200.so > gpurun_out/bench_span16.json 2> gpurun_out/bench_span16.log
for f in span32 span16; do python -c "import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', d['ms_per_step'], d['roofline']['frac'], d['device_breakdown_ms'], d['e2e']['seconds_per_step'], d['e2e']['breakdown_last_step']['decode_copy_inflate_ms'])"; done
