cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1b.log 2>&1
tail -4 gpurun_out/gpu_tests_r1b.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.log
grep -E "e2e|timing" gpurun_out/bench_r1b.log | tail -4
python -c "import json; d=json.load(open('gpurun_out/bench_r1b.json')); print(json.dumps(d['e2e'])); print(d['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
