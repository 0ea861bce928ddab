cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
$B/bamgen --out /tmp/ns.bam --contigs 906000 --reads 50000000 --seed 20260925 --median-len 4000 --sigma 0.8 --min-len 1000 --max-len 2000000 --threads 16 | tail -1
ls -la /tmp/ns.bam
for k in t1 g8; do
  echo "== $k"; CMB_INFLATE=$k CMB_PIPELINE_STATS=1 timeout 600 $B/coverm contig -m mean trimmed_mean covered_fraction -b /tmp/ns.bam -t 16 -o /dev/null --timing 2>&1 | grep -E "decode_|device_decode|#timing" | cut -c1-400
done
echo "== t1 second sample in one process (warm buffers)"; CMB_INFLATE=t1 CMB_PIPELINE_STATS=1 timeout 900 $B/coverm contig -m mean -b /tmp/ns.bam /tmp/ns.bam -t 16 -o /dev/null --timing 2>&1 | grep -E "decode_|device_decode|#timing" | cut -c1-400
