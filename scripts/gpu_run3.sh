cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "device_inflate or host_decode or in_memory or smoke or reuse" > gpurun_out/gpu_tests_persist.log 2>&1
tail -3 gpurun_out/gpu_tests_persist.log
CMB_PIPELINE_STATS=1 timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_persist.json 2> gpurun_out/bench_persist.log
grep -E "device_decode" gpurun_out/bench_persist.log | tail -3
python -c "import json; d=json.load(open('gpurun_out/bench_persist.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['seconds_per_step'], d['e2e']['step_walls_s'], d['e2e']['breakdown_last_step'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_scan_reduce -s 2 -c 1 -f -o gpurun_out/k2_full_r1e python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_k2e.log
tail -2 gpurun_out/ncu_k2e.log
