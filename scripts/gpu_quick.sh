#!/bin/bash
# Quick GPU iteration: selected tests + short bench (+ optional A/B env).  usage: bash scripts/gpu_quick.sh TAG "pytest args" "bench args"
TAG=${1:-q}; PT=${2:-tests -m gpu -x -q}; BA=${3:---steps 10 --warmup 3 --no-e2e}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest $PT > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
timeout 600 python bench.py $BA > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench.json"))
    print("value %.4g ticks/s  ms/step %.3f" % (d["value"], d["ms_per_step"]))
    print("kernels", json.dumps(d["roofline"]["all_kernels_ms_per_step"]))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "clocks", d["clocks"])
    print("time bars", json.dumps(d.get("time_bars_1min")))
    print("e2e", d.get("e2e"))
except Exception as ex:
    print("bench parse failed", ex); print(open("$O/${TAG}_bench.err").read()[-3000:])
PY
