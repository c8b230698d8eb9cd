cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29621 tests/mgpu_worker.py > gpurun_out/r1_mgpu4_parity_v21.log 2>&1; echo "parity rc=$?"; tail -6 gpurun_out/r1_mgpu4_parity_v21.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/r1_bench_4gpu_v21.json 2> gpurun_out/r1_bench_4gpu_v21.err; echo "bench rc=$?"; tail -1 gpurun_out/r1_bench_4gpu_v21.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['sweeps_per_step'], d.get('comm'), d['roofline']['frac'])"
