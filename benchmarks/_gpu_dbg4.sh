cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for env in "B2SV_PIPE_BITS=0" "B2SV_BULK=0" "B2SV_BULK=1"; do
echo "== $env"
env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 benchmarks/mgpu_trace.py --qubits 30 --steps 30 > gpurun_out/dbg4.log 2>&1
grep -E '^\{"world"' gpurun_out/dbg4.log | head -1 | cut -c1-200
grep -E "Error in|error" gpurun_out/dbg4.log | head -3 | cut -c1-300
done
