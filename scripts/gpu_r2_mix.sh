# round 2: (1) ns-size decode diagnostics after the waiting rework, (2) K1/K2 variants on config 2, (3) parity tests for them
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
$B/bamgen --out /tmp/ns.bam --contigs 906000 --reads 50000000 --seed 20260925 --median-len 4000 --sigma 0.8 --min-len 1000 --max-len 2000000 --threads 16 | tail -1
run() { echo "== $*"; env "$@" CMB_PIPELINE_STATS=1 timeout 300 $B/coverm contig -m mean trimmed_mean covered_fraction -b /tmp/ns.bam -t 16 -o /dev/null --timing 2>&1 | grep -E "decode_h2d|decode_status|device_decode" | cut -c1-330; }
run CMB_INFLATE=t1
run CMB_INFLATE=t1 CMB_T1_MAX_CTAS=592
run CMB_INFLATE=t1 CMB_PREWARM=1
run CMB_INFLATE=t1 CUDA_DEVICE_MAX_CONNECTIONS=32
run CMB_INFLATE=g8
echo "== t1 twice in one process"; CMB_INFLATE=t1 CMB_PIPELINE_STATS=1 timeout 600 $B/coverm contig -m mean -b /tmp/ns.bam /tmp/ns.bam -t 16 -o /dev/null --timing 2>&1 | grep -E "decode_h2d|decode_status|device_decode" | cut -c1-330
rm -f /tmp/ns.bam
show() { python - <<P
import json
d=json.load(open('gpurun_out/r2_bench_$1.json'))
print('$1 value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'parity',d['parity'])
print(d['device_breakdown_ms_rank0'])
P
}
for v in main k1nopre k1mb8 span64; do
  L=""; [ $v != main ] && L="--lib variants/$v.so"
  timeout 600 python bench.py --steps 20 --warmup 3 $L > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.log; echo "bench $v rc=$?"; show $v
done
# parity: default build (K1 prefetch, t1 waiting) and the SPAN=64 build
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_gpu_parity_main.log 2>&1; echo "pytest main rc=$?"; tail -3 gpurun_out/r2_gpu_parity_main.log
cp coverm_b200/libcoverm_b200.so /tmp/main.so; cp variants/span64.so coverm_b200/libcoverm_b200.so
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_gpu_parity_span64.log 2>&1; echo "pytest span64 rc=$?"; tail -3 gpurun_out/r2_gpu_parity_span64.log
cp /tmp/main.so coverm_b200/libcoverm_b200.so
