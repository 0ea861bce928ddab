# round 2, final single-GPU pass: full GPU test suite, bench lines (config 2, reference arm, ns, config 3), ncu launch list and
# --set full captures of the kernels.  Numbers printed under ncu are never bench values.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c2_final.json 2> gpurun_out/r2_bench_c2_final.log; echo "bench c2 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c2_reference.json 2> gpurun_out/r2_bench_c2_reference.log; echo "bench ref rc=$?"; cut -c1-500 gpurun_out/r2_bench_c2_reference.json
Q="python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-cold-cli --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv $Q > /dev/null 2> gpurun_out/r2_ncu_launch.log; echo "launch list rc=$?"; grep -c . gpurun_out/r2_launches.csv
for k in k2_scan_reduce k1_filter_accumulate k3_finalize; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -o gpurun_out/r2_$k -f $Q > /dev/null 2> gpurun_out/r2_ncu_$k.log; echo "ncu $k rc=$?"
done
BAM=$(ls /tmp/coverm_b200_bench/sample_c2_*.bam | head -1)
CMB_DECODE_PROFILE=1 CMB_INFLATE=t1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:kd_inflate_t1 --launch-skip 1 --launch-count 1 -o gpurun_out/r2_kd_inflate_t1 -f coverm_b200/bin/coverm contig -m mean -b $BAM -t 16 -o /dev/null > gpurun_out/r2_ncu_t1.log 2>&1; echo "ncu t1 rc=$?"
rm -f /tmp/coverm_b200_bench/sample_c2_*.bam
for cfg in ns 3; do
  timeout 1500 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg${cfg}_final.json 2> gpurun_out/r2_bench_cfg${cfg}_final.log; echo "bench $cfg rc=$?"
  rm -f /tmp/coverm_b200_bench/sample_c${cfg}_*.bam
done
python - <<'P'
import json
for n in ('c2_final','cfgns_final','cfg3_final'):
    try: d=json.load(open(f'gpurun_out/r2_bench_{n}.json'))
    except Exception as e: print(n, e); continue
    print(n,'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'cold',(d['e2e'].get('cold_cli') or {}).get('seconds'),'cpu',(d.get('cpu_baseline') or {}).get('value'),'parity',d['parity'])
P
