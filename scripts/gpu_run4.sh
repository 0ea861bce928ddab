cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_edge_cases.py tests/test_gpu_parity.py -x -q -m gpu -k "edge_cases or device_inflate or host_decode or in_memory or smoke" > gpurun_out/gpu_tests_edge.log 2>&1
tail -25 gpurun_out/gpu_tests_edge.log | cut -c1-400
CMB_PIPELINE_STATS=1 timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_crc.json 2> gpurun_out/bench_crc.log
grep -E "device_decode" gpurun_out/bench_crc.log | tail -2
python -c "import json; d=json.load(open('gpurun_out/bench_crc.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['seconds_per_step'], d['e2e']['step_walls_s'])"
