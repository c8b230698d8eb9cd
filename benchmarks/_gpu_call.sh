cd $GRAFT_REPO_ROOT
run() { python bench.py --no-cpu-baseline --steps 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],2), d['config']['sweeps_per_step'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['config']['checks'])"; }
run low5
B2SV_TILE_LOW=4 run low4
B2SV_TILE_LOW=3 run low3
B2SV_TILE_LOW=2 run low2
B2SV_TILE_LOW=3 B2SV_MAX_HEAVY=10 run low3mh10
run low5
B2SV_TILE_LOW=3 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
