# round 2: t1 inflate v2 (452 B/thread, 16 warps/SM, pipelined match copy, per-window launches), K1 fix, pair path on device
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_decode_edge_cases.py -m gpu -x -q -k "inflate or decode or edge or declined or smoke or oddities or pair" > gpurun_out/r2_gpu_tests_t1b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests_t1b.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --skip-cold-cli --skip-cpu-baseline > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.log; echo "bench $name rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_$name.json'))
print('$name value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'])
print(d['device_breakdown_ms_rank0']); print({k:v for k,v in d['e2e']['breakdown_last_step_rank0'].items() if 'decode' in k or k in ('total_s','end_sample_s')})
P
}
run t1win CMB_INFLATE=t1
run t1per CMB_INFLATE=t1 CMB_INFLATE_PERSISTENT=1
BAM=/tmp/coverm_b200_bench/sample_c2_r0_500000_10000000.bam
CMB_INFLATE_PERSISTENT=1 CMB_DECODE_PROFILE=1 CMB_PIPELINE_STATS=1 timeout 300 coverm_b200/bin/coverm contig -m mean -b $BAM -t 16 -o /dev/null 2>&1 | grep -E "decode_profile|device_decode" | head
CMB_INFLATE_PERSISTENT=1 CMB_DECODE_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_inflate_t1 --launch-skip 1 --launch-count 1 -o gpurun_out/r2_kd_inflate_t1 coverm_b200/bin/coverm contig -m mean -b $BAM -t 16 -o /dev/null > gpurun_out/r2_ncu_t1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_t1.log
ls -la gpurun_out/*.ncu-rep | tail -2
