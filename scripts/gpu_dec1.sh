set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
B=coverm_b200/bin
$B/bamgen --out /tmp/t.bam --contigs 20000 --reads 400000 --seed 5 --threads 16 > /dev/null
(CMB_PIPELINE_STATS=1 CMB_DECODE_VERIFY=1 timeout 300 $B/coverm contig -m mean trimmed_mean -b /tmp/t.bam -t 8 | md5sum) > gpurun_out/dec_small.log 2>&1
(CMB_HOST_DECODE=1 timeout 300 $B/coverm contig -m mean trimmed_mean -b /tmp/t.bam -t 8 | md5sum) >> gpurun_out/dec_small.log 2>&1
(CMB_PIPELINE_STATS=1 CMB_DECODE_VERIFY=1 timeout 300 $B/coverm contig -m mean -b tests/golden/data/1.bam -t 8 | md5sum) >> gpurun_out/dec_small.log 2>&1
(CMB_PIPELINE_STATS=1 CMB_DECODE_VERIFY=1 timeout 300 $B/coverm contig -m mean -b tests/golden/data/eg2.bam -t 8 | md5sum) >> gpurun_out/dec_small.log 2>&1
cat gpurun_out/dec_small.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "device_inflate or host_decode" > gpurun_out/dec_tests.log 2>&1
tail -5 gpurun_out/dec_tests.log
CMB_PIPELINE_STATS=1 timeout 600 python bench.py --steps 3 --warmup 1 > gpurun_out/bench_dec1.json 2> gpurun_out/bench_dec1.log
grep -E "device_decode|pipeline" gpurun_out/bench_dec1.log | tail -4
cat gpurun_out/bench_dec1.json
