run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['events_rank0_since_create'], d['config']['state_digest_rank0'])"; }
for c in 1 2 4 8; do run RG_CHUNKS=$c; done
