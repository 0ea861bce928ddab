cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.log
python -c "import json; d=json.load(open('gpurun_out/bench_r1d.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['seconds_per_step'], d['cpu_baseline']['value'], d['parity_vs_oracle_on_cpu_sample'], d['gpu_launches'], d['e2e']['gpu_launches_per_step'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1d.json 2> gpurun_out/bench_ref_r1d.log
cat gpurun_out/bench_ref_r1d.json | cut -c1-600
B="python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1d.csv $B > /dev/null 2> gpurun_out/ncu_launch.log
grep -c kd_ gpurun_out/launches_r1d.csv
