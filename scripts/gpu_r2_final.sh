# round 2, final single-GPU pass, most important first: full GPU test suite, bench lines (config 2 + reference arm), ncu launch
# list and --set full captures, north-star bench, compute-sanitizer, config 3 bench.  Numbers printed under ncu are never bench values.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
summary() { python - <<'P'
import json
for n in ('c2_final','cfgns_final','cfg3_final'):
    try: d=json.load(open(f'gpurun_out/r2_bench_{n}.json'))
    except Exception as e: print(n, 'missing'); continue
    print(n,'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'cold',(d['e2e'].get('cold_cli') or {}).get('seconds'),'cpu',(d.get('cpu_baseline') or {}).get('value'),'parity',d['parity'])
P
}
bench_cfg() { cfg=$1
  timeout 1500 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg${cfg}_final.json 2> gpurun_out/r2_bench_cfg${cfg}_final.log; echo "bench $cfg rc=$?"
  rm -f /tmp/coverm_b200_bench/sample_c${cfg}_*.bam; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c2_final.json 2> gpurun_out/r2_bench_c2_final.log; echo "bench c2 rc=$?"; summary
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c2_reference.json 2> gpurun_out/r2_bench_c2_reference.log; echo "bench ref rc=$?"; cut -c1-400 gpurun_out/r2_bench_c2_reference.json
Q="python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-cold-cli --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv $Q > /dev/null 2> gpurun_out/r2_ncu_launch.log; echo "launch list rc=$?"; grep -c . gpurun_out/r2_launches.csv
for k in k2_scan_reduce k1_filter_accumulate k3_finalize; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -o gpurun_out/r2_$k -f $Q > /dev/null 2> gpurun_out/r2_ncu_$k.log; echo "ncu $k rc=$?"
done
BAM=$(ls /tmp/coverm_b200_bench/sample_c2_*.bam | head -1)
CMB_DECODE_PROFILE=1 CMB_INFLATE=t1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:kd_inflate_t1 --launch-skip 1 --launch-count 1 -o gpurun_out/r2_kd_inflate_t1 -f coverm_b200/bin/coverm contig -m mean -b $BAM -t 16 -o /dev/null > gpurun_out/r2_ncu_t1.log 2>&1; echo "ncu t1 rc=$?"
rm -f /tmp/coverm_b200_bench/sample_c2_*.bam
bench_cfg ns; summary
# compute-sanitizer over the kernels: small synthetic BAMs; the inflate launch is serial under the tool
B=coverm_b200/bin; D=tests/golden/data
$B/bamgen --out /tmp/san.bam --contigs 300 --reads 30000 --seed 77 --median-len 3000 --min-len 200 --max-len 50000 --threads 8 > /dev/null
$B/bamgen --out /tmp/sanm.bam --contigs 200 --genomes 8 --reads 20000 --seed 78 --median-len 6000 --threads 8 > /dev/null
printf 'c0000001\ttest\tgene\t10\t900\t.\t+\t.\tID=g1\nc0000002\ttest\tgene\t1\t5000\t.\t+\t.\tID=g2\n' > /tmp/san.gff
san() { tool=$1; label=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 "$@" > /tmp/san_out.txt 2> gpurun_out/r2_sanitizer_${tool}_$label.log; rc=$?
  grep -h "ERROR SUMMARY" gpurun_out/r2_sanitizer_${tool}_$label.log /tmp/san_out.txt | tr '\n' ' '; echo " [$tool $label rc=$rc]"; }
san memcheck contig_all $B/coverm contig -m mean trimmed_mean variance covered_fraction count rpkm -b /tmp/san.bam -t 4
san memcheck pairs $B/coverm contig -m mean variance --proper-pairs-only --min-read-aligned-length-pair 250 -b /tmp/san.bam -t 4
san memcheck genome_hist $B/coverm genome -s '~' -m mean trimmed_mean variance --min-covered-fraction 0 -b /tmp/sanm.bam -t 4
san memcheck genes $B/coverm contig -m mean trimmed_mean count --gff /tmp/san.gff -b /tmp/san.bam -t 4
san memcheck filter $B/coverm filter --proper-pairs-only --min-read-aligned-length-pair 250 -b /tmp/san.bam -o /tmp/san_out.bam -t 4
san memcheck t1 env CMB_INFLATE=t1 $B/coverm contig -m mean -b /tmp/san.bam -t 4
san racecheck contig_all $B/coverm contig -m mean trimmed_mean variance covered_fraction -b /tmp/san.bam -t 4
san racecheck genome_hist $B/coverm genome -s '~' -m mean trimmed_mean variance --min-covered-fraction 0 -b /tmp/sanm.bam -t 4
bench_cfg 3; summary
