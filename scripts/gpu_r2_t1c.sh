# round 2: t1 inflate v3 (in-place prefetch, pending match, 15 warps/SM) + per-gene coverage on the device
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_genes.py tests/test_gpu_parity.py -m gpu -x -q -k "gene or inflate or declined or smoke" > gpurun_out/r2_gpu_tests_t1c.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests_t1c.log
env CMB_INFLATE=t1 timeout 600 python bench.py --steps 5 --warmup 3 --skip-cold-cli --skip-cpu-baseline > gpurun_out/r2_bench_t1c.json 2> gpurun_out/r2_bench_t1c.log; echo "bench rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_t1c.json'))
print('t1c value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'])
print({k:v for k,v in d['e2e']['breakdown_last_step_rank0'].items() if 'decode' in k or k in ('total_s','end_sample_s')})
P
BAM=/tmp/coverm_b200_bench/sample_c2_r0_500000_10000000.bam
CMB_DECODE_PROFILE=1 CMB_PIPELINE_STATS=1 timeout 300 coverm_b200/bin/coverm contig -m mean -b $BAM -t 16 -o /dev/null 2>&1 | grep -E "decode_profile|device_decode" | head
CMB_DECODE_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_inflate_t1 --launch-skip 1 --launch-count 1 -o gpurun_out/r2_kd_inflate_t1c coverm_b200/bin/coverm contig -m mean -b $BAM -t 16 -o /dev/null > gpurun_out/r2_ncu_t1c.log 2>&1; echo "ncu rc=$?"
