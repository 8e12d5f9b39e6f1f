#!/bin/bash
# One gpurun call: GPU tests of the current build, then A/B of the elementwise load batching (AWR_EW_UNROLL) and the head split (AWR_HEAD_SPLIT).
set -u
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/ab_t0
timeout 420 python -m pytest tests -m gpu -x -q > $O/ab_pytest.log 2>&1; echo "pytest exit $?" >> $O/ab_pytest.log
tail -3 $O/ab_pytest.log
timeout 200 python bench.py --steps 100 --cpu-steps 0 --layers ab_layers_new.md > $O/ab_bench_new.json 2> $O/ab_bench_new.err; echo "bench new $?"
AWR_EW_UNROLL=1 AWR_HEAD_SPLIT=1 timeout 200 python bench.py --steps 100 --cpu-steps 0 --no-parity --layers ab_layers_old.md > $O/ab_bench_old.json 2> $O/ab_bench_old.err; echo "bench old $?"
for s in 1 2 4; do AWR_HEAD_SPLIT=$s timeout 100 python tools/bench_head.py > $O/ab_head_s$s.log 2>&1; echo "head s$s $?"; done
AWR_EW_UNROLL=2 timeout 200 python bench.py --steps 100 --cpu-steps 0 --no-parity > $O/ab_bench_u2.json 2> $O/ab_bench_u2.err; echo "bench u2 $?"
timeout 200 python bench.py --steps 50 --cpu-steps 0 --no-parity --gpu-eager-baseline > $O/ab_bench_eager.json 2> $O/ab_bench_eager.err; echo "bench eager $?"
date +%s > $O/ab_t1
for f in new old u2 eager; do python - <<PY
import json
try:
    d = json.loads(open("$O/ab_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline_head"]["pair_us"], d.get("parity"), d.get("gpu_eager_baseline"))
except Exception as e:
    print("$f", "ERR", e)
PY
done
cat $O/ab_head_s*.log | grep "B32"
