# round 2: compute-sanitizer memcheck + racecheck over the kernels (fixtures and a small synthetic BAM), logs -> gpurun_out/
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
D=tests/golden/data
$B/bamgen --out /tmp/san.bam --contigs 300 --reads 30000 --seed 77 --median-len 3000 --min-len 200 --max-len 50000 --threads 8 > /dev/null
$B/bamgen --out /tmp/sanm.bam --contigs 200 --genomes 8 --reads 20000 --seed 78 --median-len 6000 --threads 8 > /dev/null
printf 'c0000001\ttest\tgene\t10\t900\t.\t+\t.\tID=g1\nc0000002\ttest\tgene\t1\t5000\t.\t+\t.\tID=g2\nc0000002\ttest\tgene\t100\t300\t.\t-\t.\tID=g3\n' > /tmp/san.gff
run() { # tool, label, args...
  tool=$1; label=$2; shift 2
  CMB_INFLATE_WINDOWS=1 timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 "$@" > /tmp/san_out.txt 2> gpurun_out/r2_sanitizer_${tool}_$label.log
  rc=$?
  tail -3 gpurun_out/r2_sanitizer_${tool}_$label.log | tr '\n' ' '; echo " [$tool $label rc=$rc]"
}
for tool in memcheck racecheck; do
  run $tool contig_all $B/coverm contig -m mean trimmed_mean variance covered_fraction count rpkm -b /tmp/san.bam -t 4
  run $tool fixture_1bam $B/coverm contig -m mean trimmed_mean variance -b $D/1.bam -t 4
  run $tool pairs $B/coverm contig -m mean variance --proper-pairs-only --min-read-aligned-length-pair 250 -b /tmp/san.bam -t 4
  run $tool genome_hist $B/coverm genome -s '~' -m mean trimmed_mean variance --min-covered-fraction 0 -b /tmp/sanm.bam -t 4
  run $tool genes $B/coverm contig -m mean trimmed_mean count --gff /tmp/san.gff -b /tmp/san.bam -t 4
  run $tool filter $B/coverm filter --proper-pairs-only --min-read-aligned-length-pair 250 -b /tmp/san.bam -o /tmp/san_out.bam -t 4
  run $tool hostdecode env CMB_HOST_DECODE=1 $B/coverm contig -m mean trimmed_mean -b /tmp/san.bam -t 4
  run $tool persistent_g8 env CMB_INFLATE=g8 CMB_INFLATE_WINDOWS=0 $B/coverm contig -m mean -b $D/7seqs.reads_for_seq1_and_seq2.bam -t 4
done
grep -l "ERROR SUMMARY: 0 errors" gpurun_out/r2_sanitizer_*.log | wc -l; ls gpurun_out/r2_sanitizer_*.log | wc -l
