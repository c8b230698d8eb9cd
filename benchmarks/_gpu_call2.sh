cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py > gpurun_out/r1_mgpu2_parity_v20.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r1_mgpu2_parity_v20.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1_bench_2gpu_v20.json 2> gpurun_out/r1_bench_2gpu_v20.err; echo "bench rc=$?"; tail -1 gpurun_out/r1_bench_2gpu_v20.json | cut -c1-400; tail -1 gpurun_out/r1_bench_2gpu_v20.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['sweeps_per_step'], d.get('comm'), d['roofline']['frac'])"
