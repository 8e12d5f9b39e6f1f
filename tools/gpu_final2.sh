#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > $O/fin2_pytest.log 2>&1; echo "pytest exit $?" >> $O/fin2_pytest.log; tail -5 $O/fin2_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/fin2_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/fin2_smoke.log
timeout 400 python bench.py > $O/fin2_bench.json 2> $O/fin2_bench.err; echo "bench $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/fin2_bench_ref.json 2> $O/fin2_bench_ref.err; echo "bench ref $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/fin2_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["roofline_head"], d["cpu_baseline"], d["parity"]["uvd_max_abs_diff"], d["clocks"])
print(open("gpurun_out/fin2_bench_ref.json").read()[:300])
PY
