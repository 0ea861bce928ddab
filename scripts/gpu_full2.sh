cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1d.log 2>&1
tail -6 gpurun_out/gpu_tests_r1d.log | cut -c1-300
