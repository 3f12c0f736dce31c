run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['events_rank0_since_create'])"; }
run RG_PF_WPB=16
run RG_PF_WPB=8
run RG_PF_WPB=4
run RG_PF_WPB=16 RG_PREFETCH_EVERY=2
run RG_PF_WPB=8 RG_PREFETCH_EVERY=2
run RG_PF_WPB=8 RG_PREFETCH_EVERY=4
