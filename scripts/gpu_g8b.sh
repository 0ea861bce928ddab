cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
$B/bamgen --out /tmp/m2.bam --contigs 100000 --reads 2200000 --seed 9 --threads 16 > /dev/null
for g in 1 0; do
  echo "== G8=$g"
  (CMB_INFLATE_G8=$g CMB_PIPELINE_STATS=1 CMB_DECODE_PROFILE=1 timeout 120 $B/coverm contig -m mean trimmed_mean -b /tmp/m2.bam -t 16 | md5sum) 2>&1 | grep -v "^#pipeline"
done
CMB_DECODE_PROFILE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:kd_inflate_g8 -s 1 -c 1 -f -o gpurun_out/kd_inflate_g8_a $B/coverm contig -m mean -b /tmp/m2.bam -t 16 > /dev/null 2> gpurun_out/ncu_g8.log
tail -2 gpurun_out/ncu_g8.log
for g in 1 0; do
CMB_INFLATE_G8=$g CMB_PIPELINE_STATS=1 timeout 400 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_g8_$g.json 2> gpurun_out/bench_g8_$g.log
grep -E "device_decode" gpurun_out/bench_g8_$g.log | tail -2
python -c "import json; d=json.load(open('gpurun_out/bench_g8_$g.json')); print('G8=$g', d['e2e']['seconds_per_step'], d['e2e']['step_walls_s'])"
done
