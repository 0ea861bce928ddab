# round 2, after K2's static schedule became the default: full GPU suite again, config 2 bench line, K2 capture, launch list,
# memcheck with the tool's reports kept
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_final2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_final2.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c2_final2.json 2> gpurun_out/r2_bench_c2_final2.log; echo "bench c2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_c2_final2.json'))
print('c2_final2 value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_step'],'cold',(d['e2e'].get('cold_cli') or {}).get('seconds'),'cpu',(d.get('cpu_baseline') or {}).get('value'),'parity',d['parity'])
print(d['device_breakdown_ms_rank0'])
P
Q="python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-cold-cli --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv $Q > /dev/null 2> gpurun_out/r2_ncu_launch.log; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_scan_reduce --launch-skip 2 --launch-count 1 -o gpurun_out/r2_k2_scan_reduce -f $Q > /dev/null 2> gpurun_out/r2_ncu_k2_scan_reduce.log; echo "ncu k2 rc=$?"
rm -f /tmp/coverm_b200_bench/sample_c2_*.bam
B=coverm_b200/bin
$B/bamgen --out /tmp/san.bam --contigs 300 --reads 30000 --seed 77 --median-len 3000 --min-len 200 --max-len 50000 --threads 8 > /dev/null
$B/bamgen --out /tmp/sanm.bam --contigs 200 --genomes 8 --reads 20000 --seed 78 --median-len 6000 --threads 8 > /dev/null
printf 'c0000001\ttest\tgene\t10\t900\t.\t+\t.\tID=g1\nc0000002\ttest\tgene\t1\t5000\t.\t+\t.\tID=g2\n' > /tmp/san.gff
san() { tool=$1; label=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/r2_sanitizer_${tool}_$label.log "$@" > /dev/null 2> /tmp/san_err.txt; rc=$?
  grep -hE "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_${tool}_$label.log | tr '\n' ' '; echo " [$tool $label rc=$rc]"; }
san memcheck contig_all $B/coverm contig -m mean trimmed_mean variance covered_fraction count rpkm -b /tmp/san.bam -t 4
san memcheck pairs $B/coverm contig -m mean variance --proper-pairs-only --min-read-aligned-length-pair 250 -b /tmp/san.bam -t 4
san memcheck genome_hist $B/coverm genome -s '~' -m mean trimmed_mean variance --min-covered-fraction 0 -b /tmp/sanm.bam -t 4
san memcheck genes $B/coverm contig -m mean trimmed_mean count --gff /tmp/san.gff -b /tmp/san.bam -t 4
san memcheck filter $B/coverm filter --proper-pairs-only --min-read-aligned-length-pair 250 -b /tmp/san.bam -o /tmp/san_out.bam -t 4
san memcheck t1 env CMB_INFLATE=t1 $B/coverm contig -m mean -b /tmp/san.bam -t 4
san racecheck contig_all $B/coverm contig -m mean trimmed_mean variance covered_fraction -b /tmp/san.bam -t 4
