run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 600 --warmup 100 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['events_rank0_since_create'])"; }
run RG_PREFETCH_EVERY=1
run RG_PREFETCH_EVERY=4
run RG_PREFETCH_EVERY=16
run RG_PREFETCH_WARPS=592
run RG_PREFETCH_WARPS=592 RG_PREFETCH_EVERY=4
run RG_PREFETCH_WARPS=296 RG_PREFETCH_EVERY=8
run RG_PREFETCH_WARPS=592 RG_BG_PRIO=h
