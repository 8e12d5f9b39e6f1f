#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 150 python -m pytest tests/test_bn_gpu.py tests/test_fused_stem_gpu.py -x -q --timeout 60 > $O/ab3_pytest_bn.log 2>&1; echo "pytest-bn exit $?" >> $O/ab3_pytest_bn.log; tail -15 $O/ab3_pytest_bn.log
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > $O/ab3_pytest.log 2>&1; echo "pytest exit $?" >> $O/ab3_pytest.log; tail -6 $O/ab3_pytest.log
timeout 200 python bench.py --steps 100 --cpu-steps 0 --layers ab3_layers_new.md > $O/ab3_bench_new.json 2> $O/ab3_bench_new.err; echo "bench new $?"
AWR_BN_FUSED=0 timeout 200 python bench.py --steps 100 --cpu-steps 0 --no-parity > $O/ab3_bench_nofused.json 2> $O/ab3_bench_nofused.err; echo "bench nofused $?"
for f in new nofused; do python - <<PY
import json
try:
    d = json.loads(open("$O/ab3_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["launches_per_step"], (d.get("parity") or {}).get("uvd_max_abs_diff"))
    print({k: v["ms_per_step"] for k, v in d["kernel_classes"].items()})
except Exception as e:
    print("$f", "ERR", e)
PY
done
tail -3 $O/ab3_bench_new.err
