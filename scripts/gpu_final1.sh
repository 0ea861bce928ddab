cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1c.log 2>&1
tail -4 gpurun_out/gpu_tests_r1c.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.log
python -c "import json; d=json.load(open('gpurun_out/bench_r1c.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['e2e']['value'], d['e2e']['seconds_per_step'], d['cpu_baseline']['value'], d['parity_vs_oracle_on_cpu_sample'])"
B="python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv $B > /dev/null 2> gpurun_out/ncu_launch.log
for k in k2_scan k3_final k1_filter; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${k}_r1c $B > /dev/null 2> gpurun_out/ncu_$k.log
done
coverm_b200/bin/bamgen --out /tmp/m.bam --contigs 50000 --reads 1000000 --seed 7 --threads 16 > /dev/null
CMB_DECODE_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_inflate -s 1 -c 1 -f -o gpurun_out/kd_inflate_r1c coverm_b200/bin/coverm contig -m mean -b /tmp/m.bam -t 16 > /dev/null 2> gpurun_out/ncu_kd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_extract -c 1 -f -o gpurun_out/kd_extract_r1c coverm_b200/bin/coverm contig -m mean -b /tmp/m.bam -t 16 > /dev/null 2>> gpurun_out/ncu_kd.log
ls -la gpurun_out/*.ncu-rep | tail -8
