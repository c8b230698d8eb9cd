# Late round 2: re-validation of the shipped tile kernel (16-byte c64 loader + FFMA2) on one B200.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_late_pytest_gpu.log
cat gpurun_out/r2_late_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_late_bench_1gpu.json 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-adjoint --dtype c64 --qubits 31 > gpurun_out/r2_late_bench_c64.json 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-adjoint --workload controlled --layers 2 > gpurun_out/r2_late_bench_controlled.json 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-adjoint --workload controlled --layers 2 --dtype c64 --qubits 31 > gpurun_out/r2_late_bench_controlled_c64.json 2>/dev/null
for f in 1gpu c64 controlled controlled_c64; do python -c "
import json
d=json.loads(open('gpurun_out/r2_late_bench_$f.json').readline()); r=d['roofline']
print('$f', 'ms/step', round(d['ms_per_step'],2), 'passes', d['schedule']['sweeps_per_step'], 'ms/pass', round(r['avg_launch_ms'],3), 'frac', round(r['frac'],3), 'parity', d.get('parity'), 'e2e', d['e2e']['value'] if 'e2e' in d else None)"; done
