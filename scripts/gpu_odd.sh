cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_decode_edge_cases.py -x -q -m gpu -k "oddities" > gpurun_out/gpu_tests_odd.log 2>&1
tail -12 gpurun_out/gpu_tests_odd.log | cut -c1-500
