run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config']['events_rank0_since_create'], d['config']['state_digest_rank0'])"; }
run A=1
run ROGUE_B200_LIB=$PWD/rogue-gym_b200/variants/lib_noinl.so
ROGUE_B200_LIB=$PWD/rogue-gym_b200/variants/lib_noinl.so timeout 300 python tools/bench_reset.py 2>&1 | cut -c1-200
