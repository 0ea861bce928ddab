# round 2: the north-star workload (5 Gbp / 50 M reads, contig) and configs[2] (genome, 1000 MAGs / 50 M reads) at N=1, with
# full-file parity against the oracle, the CPU baseline on the same file and a cold CLI run; plus a t1 concurrency experiment
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BAM=/tmp/coverm_b200_bench/sample_c2_r0_500000_10000000.bam
timeout 300 python bench.py --steps 2 --warmup 1 --skip-cold-cli --skip-cpu-baseline > /dev/null 2>&1   # generates the config-2 file
for cap in 740 370 185; do
  CMB_INFLATE=t1 CMB_T1_MAX_CTAS=$cap CMB_DECODE_PROFILE=1 timeout 300 coverm_b200/bin/coverm contig -m mean -b $BAM -t 16 -o /dev/null 2>&1 | grep -E "decode_profile" | sed "s/^/cap=$cap /" | cut -c1-160
done
for cfg in ns 3; do
  timeout 1500 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg$cfg.json 2> gpurun_out/r2_bench_cfg$cfg.log; echo "bench $cfg rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_cfg$cfg.json'))
print('$cfg', d['config']['workload'])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'cold',d['e2e']['cold_cli']['seconds'] if d['e2e']['cold_cli'] else None,'cpu',d['cpu_baseline']['value'], d['cpu_baseline']['seconds_per_run'],'parity',d['parity'])
print(d['device_breakdown_ms_rank0']); print({k:v for k,v in d['e2e']['breakdown_last_step_rank0'].items() if 'decode' in k or k in ('total_s','end_sample_s')})
P
  rm -f /tmp/coverm_b200_bench/sample_c${cfg}_*.bam
done
