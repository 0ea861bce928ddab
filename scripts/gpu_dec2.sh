set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
B=coverm_b200/bin
$B/bamgen --out /tmp/m.bam --contigs 50000 --reads 1000000 --seed 7 --threads 16 > /dev/null
for i in 1 2; do (CMB_PIPELINE_STATS=1 CMB_DECODE_PROFILE=1 timeout 300 $B/coverm contig -m mean trimmed_mean -b /tmp/m.bam /tmp/m.bam -t 16 | md5sum) > gpurun_out/dec_prof.log 2>&1; done
cat gpurun_out/dec_prof.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "device_inflate or host_decode" > gpurun_out/dec_tests.log 2>&1
tail -5 gpurun_out/dec_tests.log
CMB_DECODE_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_inflate -c 8 -o gpurun_out/kd_inflate_r1 -f $B/coverm contig -m mean -b /tmp/m.bam -t 16 > /dev/null 2> gpurun_out/ncu_inflate.log
tail -3 gpurun_out/ncu_inflate.log
