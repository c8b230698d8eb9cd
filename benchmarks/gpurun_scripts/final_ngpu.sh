set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_worker.py > gpurun_out/r2_final_worker$N.log 2>&1
tail -4 gpurun_out/r2_final_worker$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_${N}gpu.raw 2> gpurun_out/r2_final_bench_${N}gpu.err
grep '^{"metric"' gpurun_out/r2_final_bench_${N}gpu.raw > gpurun_out/r2_final_bench_${N}gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2_final_bench_${N}gpu.json"))
print("N=$N ms/step", d["ms_per_step"], "value", d["value"], "sweeps", d["schedule"]["sweeps_per_step"], "parity", d["parity"]["ok"], d["parity"].get("cases"))
print(d["comm"]); print(d.get("roofline_nvlink")); print(d.get("roofline_step")); print(d.get("adjoint_jacobian")); print(d["clocks"])
PY
tail -3 gpurun_out/r2_final_bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 benchmarks/mgpu_trace.py --qubits 30 --out gpurun_out/r2_final_trace$N.json > gpurun_out/r2_final_trace$N.log 2>&1
if [ "$N" = "8" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 8 --qubits 33 --layers 2 --steps 3 --warmup 1 --no-parity --no-adjoint > gpurun_out/r2_final_bench_8gpu_36q.raw 2> gpurun_out/r2_final_bench_8gpu_36q.err
  grep '^{"metric"' gpurun_out/r2_final_bench_8gpu_36q.raw > gpurun_out/r2_final_bench_8gpu_36q.json
  python -c "
import json
d=json.load(open('gpurun_out/r2_final_bench_8gpu_36q.json'))
print('36q ms/step', d['ms_per_step'], d['schedule'], d['comm'], d.get('roofline_step'))"
fi
