# round 2: validate kd_inflate_t1 (thread-per-block inflate) + K2 unit-mask; A/B of the inflate kernels on config 2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_decode_edge_cases.py -m gpu -x -q -k "inflate or decode or edge or declined or smoke or oddities" > gpurun_out/r2_gpu_tests_t1.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests_t1.log
for K in t1 g8; do
  CMB_INFLATE=$K timeout 600 python bench.py --steps 5 --warmup 3 --skip-cold-cli --skip-cpu-baseline > gpurun_out/r2_bench_$K.json 2> gpurun_out/r2_bench_$K.log; echo "bench $K rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_$K.json'))
print('$K value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'])
print(d['device_breakdown_ms_rank0']); print({k:v for k,v in d['e2e']['breakdown_last_step_rank0'].items() if 'decode' in k or k in ('total_s','end_sample_s')})
P
done
CMB_DECODE_PROFILE=1 CMB_PIPELINE_STATS=1 timeout 300 coverm_b200/bin/coverm contig -m mean -b /tmp/coverm_b200_bench/sample_c2_r0_500000_10000000.bam -t 16 -o /dev/null 2>&1 | grep -E "decode_profile|device_decode" | head
