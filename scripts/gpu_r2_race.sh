# round 2: compute-sanitizer racecheck (shared-memory hazards) with the tool's own report kept
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
$B/bamgen --out /tmp/san.bam --contigs 300 --reads 30000 --seed 77 --median-len 3000 --min-len 200 --max-len 50000 --threads 8 > /dev/null
$B/bamgen --out /tmp/sanm.bam --contigs 200 --genomes 8 --reads 20000 --seed 78 --median-len 6000 --threads 8 > /dev/null
race() { label=$1; shift
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 --log-file gpurun_out/r2_sanitizer_racecheck_$label.log "$@" > /dev/null 2> /tmp/race_err.txt; rc=$?
  grep -h "RACECHECK SUMMARY" gpurun_out/r2_sanitizer_racecheck_$label.log | tr '\n' ' '; echo " [racecheck $label rc=$rc]"
  grep -hE "hazard detected|Race reported" gpurun_out/r2_sanitizer_racecheck_$label.log | sed -E 's/0x[0-9a-f]+/X/g; s/thread \([0-9,]+\)/thread T/g; s/block \([0-9,]+\)/block B/g' | sort | uniq -c | sort -rn | head -12 | cut -c1-260; }
race contig_all $B/coverm contig -m mean trimmed_mean variance covered_fraction -b /tmp/san.bam -t 4
race genome_hist $B/coverm genome -s '~' -m mean trimmed_mean variance --min-covered-fraction 0 -b /tmp/sanm.bam -t 4
race pairs_t1 env CMB_INFLATE=t1 $B/coverm contig -m mean --proper-pairs-only -b /tmp/san.bam -t 4
