# round 2: t1 first round dealt out column-wise (a warp's lanes hold blocks spread over the file) vs in launch order
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inflate or declined or retries or smoke or memory_is_short or c_abi" > gpurun_out/r2_gpu_tests_mix4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests_mix4.log
gen() { $B/bamgen --out /tmp/$1.bam --contigs $2 --reads $3 --seed 20260924 --median-len 4000 --sigma 0.8 --min-len 1000 --max-len 2000000 --threads 16 > /dev/null; }
run() { f=$1; shift; echo "== $f $*"; env "$@" CMB_PIPELINE_STATS=1 timeout 300 $B/coverm contig -m mean trimmed_mean covered_fraction -b /tmp/$f.bam -t 16 -o /dev/null --timing 2>&1 | grep -E "decode_profile|decode_status|device_decode|ERROR" | cut -c1-300; }
gen c2 500000 10000000
run c2 CMB_X=1
run c2 CMB_T1_IN_ORDER=1
run c2 CMB_T1_LANES=32
run c2 CMB_DECODE_PROFILE=1
gen half 250000 5000000; run half CMB_T1_MIN_BLOCKS=0; run half CMB_INFLATE=g8
gen eighth 62500 1250000; run eighth CMB_T1_MIN_BLOCKS=0; run eighth CMB_INFLATE=g8
rm -f /tmp/c2.bam /tmp/half.bam /tmp/eighth.bam
show() { python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_$1.json'))
print('$1 value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'cold',(d['e2e'].get('cold_cli') or {}).get('seconds'),'parity',d['parity'])
print([round(x,3) for x in d['e2e']['step_walls_s']])
P
grep "e2e per step" gpurun_out/r2_bench_$1.log | cut -c1-300; grep "host timing" gpurun_out/r2_bench_$1.log | cut -c1-420; }
timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_c2d.json 2> gpurun_out/r2_bench_c2d.log; echo "bench c2 rc=$?"; show c2d
timeout 1500 python bench.py --config ns --steps 3 --warmup 3 --skip-cpu-baseline --skip-cold-cli > gpurun_out/r2_bench_cfgnsd.json 2> gpurun_out/r2_bench_cfgnsd.log; echo "bench ns rc=$?"; show cfgnsd
