cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMB_PIPELINE_STATS=1 timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_dec2.json 2> gpurun_out/bench_dec2.log
grep -E "device_decode|e2e|timing" gpurun_out/bench_dec2.log | tail -12
cat gpurun_out/bench_dec2.json | python -c "import json,sys; d=json.load(sys.stdin); print(json.dumps(d['e2e'])); print(d['value'], d['cpu_baseline'])"
