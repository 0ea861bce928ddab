cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
$B/bamgen --out /tmp/big.bam --contigs 1000000 --reads 1000000 --seed 41 --threads 16 | tail -1
timeout 200 $B/coverm contig -m mean trimmed_mean covered_fraction variance -b /tmp/big.bam -t 16 --timing > /tmp/gpu.out 2> /tmp/gpu.err; echo "gpu rc=$?"; tail -2 /tmp/gpu.err | cut -c1-300
timeout 300 oracle/coverm_oracle contig -m mean trimmed_mean covered_fraction variance -b /tmp/big.bam -t 16 > /tmp/or.out 2> /tmp/or.err; echo "oracle rc=$?"
md5sum /tmp/gpu.out /tmp/or.out | awk '{print $1}' | uniq -c; wc -l /tmp/gpu.out
