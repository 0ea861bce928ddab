# round 2: where does the g8 "misaligned address" come from (memcheck), t1 lanes-per-warp experiment, inflate tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=coverm_b200/bin
D=tests/golden/data
for mode in "CMB_INFLATE=g8" "CMB_INFLATE=w1" "CMB_INFLATE=t1"; do
  echo "== memcheck $mode"; env $mode timeout 300 compute-sanitizer --tool memcheck --print-limit 3 $B/coverm contig -m mean -b $D/2seqs.reads_for_seq1.bam 2>&1 | grep -vE "^=========\s*$" | head -40 | cut -c1-220
done
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inflate or declined or retries or smoke or memory_is_short or c_abi" > gpurun_out/r2_gpu_tests_mix3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gpu_tests_mix3.log
gen() { $B/bamgen --out /tmp/$1.bam --contigs $2 --reads $3 --seed 20260924 --median-len 4000 --sigma 0.8 --min-len 1000 --max-len 2000000 --threads 16 > /dev/null; }
run() { f=$1; shift; echo "== $f $*"; env "$@" CMB_PIPELINE_STATS=1 timeout 300 $B/coverm contig -m mean trimmed_mean covered_fraction -b /tmp/$f.bam -t 16 -o /dev/null --timing 2>&1 | grep -E "decode_profile|decode_status|device_decode|ERROR" | cut -c1-300; }
gen c2 500000 10000000
run c2 CMB_DECODE_PROFILE=1
run c2 CMB_DECODE_PROFILE=1 CMB_T1_LANES=32
run c2 CMB_DECODE_PROFILE=1 CMB_T1_LANES=16
run c2 CMB_DECODE_PROFILE=1 CMB_T1_LANES=11
run c2 CMB_INFLATE=g8
gen half 250000 5000000; run half CMB_T1_MIN_BLOCKS=0; run half CMB_T1_MIN_BLOCKS=0 CMB_T1_LANES=32; run half CMB_INFLATE=g8
gen eighth 62500 1250000; run eighth CMB_T1_MIN_BLOCKS=0; run eighth CMB_T1_MIN_BLOCKS=0 CMB_T1_LANES=32; run eighth CMB_INFLATE=g8
rm -f /tmp/c2.bam /tmp/half.bam /tmp/eighth.bam
