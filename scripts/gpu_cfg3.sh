cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
$B/bamgen --out /tmp/g.bam --contigs 150000 --genomes 1000 --reads 5000000 --seed 31 --median-len 15000 --definition-out /tmp/g.tsv --threads 16 | tail -1
date +%s.%N
for args in "genome -s ~ -m mean trimmed_mean covered_fraction --min-read-percent-identity 95" "genome --genome-definition /tmp/g.tsv -m relative_abundance mean variance covered_bases count" "contig -m mean trimmed_mean variance rpkm tpm covered_fraction"; do
  timeout 300 $B/coverm $args -b /tmp/g.bam -t 16 --timing > /tmp/gpu.out 2> /tmp/gpu.err; tail -3 /tmp/gpu.err; date +%s.%N
  timeout 600 oracle/coverm_oracle $args -b /tmp/g.bam -t 16 > /tmp/or.out 2> /tmp/or.err; tail -1 /tmp/or.err; date +%s.%N
  md5sum /tmp/gpu.out /tmp/or.out | awk '{print $1}' | uniq -c
done
