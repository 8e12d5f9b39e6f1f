#!/bin/bash
# One gpurun call that re-validates a build on a B200: GPU tests, smoke, the default bench line, and (optionally) A/B runs of the
# library's tuning switches.  Usage:  gpurun --timeout 600 -- 'bash tools/gpu_check.sh [ab]'
#   switches (read once per process by libawr_b200.so / engine.py):
#     AWR_EW_UNROLL=2|4        16-byte loads in flight per thread and tensor in the BatchNorm passes (default 4)
#     AWR_HEAD_SPLIT=2|4       CTAs per (frame, joint) pair in the fused head kernels (default 1)
#     AWR_BN_FUSED=1           single-launch BatchNorm backward for small tensors (default off: measured slower)
#     AWR_STEM_TC=0            stem conv / weight gradient on the CUDA-core kernels instead of tcgen05
#     AWR_B200_OPT_OVERLAP=0   one optimizer launch after backward instead of one per gradient bucket
#     AWR_B200_POOL_PASS0=full full-resolution pass 0 of the stem BatchNorm backward
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > $O/check_pytest.log 2>&1; echo "pytest exit $?" >> $O/check_pytest.log; tail -4 $O/check_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/check_smoke.log 2>&1; echo "smoke exit $?"
timeout 300 python bench.py --layers check_layers.md > $O/check_bench.json 2> $O/check_bench.err; echo "bench exit $?"
if [ "${1:-}" = "ab" ]; then
  for kv in AWR_EW_UNROLL=2 AWR_HEAD_SPLIT=2 AWR_BN_FUSED=1 AWR_STEM_TC=0 AWR_B200_OPT_OVERLAP=0 AWR_B200_POOL_PASS0=full; do
    env $kv timeout 200 python bench.py --steps 100 --cpu-steps 0 --no-parity --no-gpu-eager-baseline > $O/check_bench_${kv%%=*}.json 2> /dev/null; echo "$kv exit $?"
  done
fi
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/check_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv frac", d["roofline"]["frac"], "head frac", d["roofline_head"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
