cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "second_device_pass or memory_is_short or device_inflate" > gpurun_out/gpu_tests_retry.log 2>&1
tail -8 gpurun_out/gpu_tests_retry.log | cut -c1-400
