run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['events_rank0_since_create'])"; }
run RG_PLAYER_BLOCKS=0
run RG_PLAYER_BLOCKS=4736
run RG_PLAYER_BLOCKS=9472
run RG_PLAYER_BLOCKS=18944
