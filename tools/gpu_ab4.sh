#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > $O/ab4_pytest.log 2>&1; echo "pytest exit $?" >> $O/ab4_pytest.log; tail -6 $O/ab4_pytest.log
timeout 200 python bench.py --steps 100 --cpu-steps 0 --layers ab4_layers_new.md > $O/ab4_bench_new.json 2> $O/ab4_bench_new.err; echo "bench new $?"
AWR_EW_UNROLL=2 timeout 200 python bench.py --steps 100 --cpu-steps 0 --no-parity > $O/ab4_bench_u2.json 2> $O/ab4_bench_u2.err; echo "bench u2 $?"
AWR_B200_POOL_PASS0=full timeout 200 python bench.py --steps 100 --cpu-steps 0 --no-parity > $O/ab4_bench_poolfull.json 2> $O/ab4_bench_poolfull.err; echo "bench poolfull $?"
for f in new u2 poolfull; do python - <<PY
import json
try:
    d = json.loads(open("$O/ab4_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["launches_per_step"], (d.get("parity") or {}).get("uvd_max_abs_diff"), d["loss"])
    print({k: v["ms_per_step"] for k, v in d["kernel_classes"].items()})
except Exception as e:
    print("$f", "ERR", e)
PY
done
tail -3 $O/ab4_bench_new.err
