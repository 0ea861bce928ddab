cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_r1d.json 2> gpurun_out/bench_n2_r1d.log
tail -3 gpurun_out/bench_n2_r1d.log | cut -c1-300
python -c "import json; d=json.loads(open('gpurun_out/bench_n2_r1d.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['e2e']['seconds_per_step'], d['device_breakdown_ms'])"
