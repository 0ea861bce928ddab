cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
$B/bamgen --out /tmp/m.bam --contigs 50000 --reads 1000000 --seed 7 --threads 16 > /dev/null
for g in 1 0; do
  echo "== G8=$g"
  (CMB_INFLATE_G8=$g CMB_PIPELINE_STATS=1 CMB_DECODE_VERIFY=1 CMB_DECODE_PROFILE=1 timeout 120 $B/coverm contig -m mean trimmed_mean -b /tmp/m.bam -t 16 | md5sum) 2>&1 | grep -v "^#pipeline"
done
for f in 1.bam eg2.bam tpm_test.bam; do (CMB_PIPELINE_STATS=1 CMB_DECODE_VERIFY=1 timeout 120 $B/coverm contig -m mean -b tests/golden/data/$f -t 8 | md5sum) 2>&1 | grep -E "decode_|device_decode|-$"; done
timeout 600 python -m pytest tests/test_decode_edge_cases.py tests/test_gpu_parity.py -x -q -m gpu -k "edge_cases or device_inflate" > gpurun_out/gpu_tests_g8.log 2>&1
tail -12 gpurun_out/gpu_tests_g8.log | cut -c1-300
