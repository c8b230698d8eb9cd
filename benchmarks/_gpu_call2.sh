cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py > gpurun_out/r1_mgpu2_parity_v21.log 2>&1; echo "parity rc=$?"; tail -12 gpurun_out/r1_mgpu2_parity_v21.log
timeout 300 python -m pytest tests -m gpu -x -q -k "probs or sampling" 2>&1 | tail -3
