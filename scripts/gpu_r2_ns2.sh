# round 2: ns + cfg3 again after the polling back-off; config 2 for the outlier check
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_$1.json'))
print('$1', d['config']['workload'][:90])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'cold',d['e2e']['cold_cli']['seconds'] if d['e2e'].get('cold_cli') else None,'cpu',(d['cpu_baseline'] or {}).get('value'),'parity',d['parity'])
print([round(x,3) for x in d['e2e']['step_walls_s']])
P
grep "e2e per step" gpurun_out/r2_bench_$1.log | cut -c1-400; grep "host timing" gpurun_out/r2_bench_$1.log | cut -c1-500; }
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c2b.json 2> gpurun_out/r2_bench_c2b.log; echo "bench c2 rc=$?"; show c2b
for cfg in ns 3; do
  timeout 1500 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg${cfg}b.json 2> gpurun_out/r2_bench_cfg${cfg}b.log; echo "bench $cfg rc=$?"; show cfg${cfg}b
  rm -f /tmp/coverm_b200_bench/sample_c${cfg}_*.bam
done
