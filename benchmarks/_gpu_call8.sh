cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_worker.py > gpurun_out/r1_mgpu8_parity_v23.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r1_mgpu8_parity_v23.log | cut -c1-200 )
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r1_bench_8gpu_v23.json 2> gpurun_out/r1_bench_8gpu_v23.err; echo "bench rc=$?"; tail -1 gpurun_out/r1_bench_8gpu_v23.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['sweeps_per_step'], d.get('comm'), d['roofline']['frac'])"
tail -3 gpurun_out/r1_bench_8gpu_v23.err | cut -c1-300
