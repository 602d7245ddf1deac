#!/bin/bash
# round 2, call 16 (1 GPU): full GPU suite + full bench line (legs included) on the current tree
mkdir -p gpurun_out; P=gpurun_out/c16
timeout 1500 python -m pytest tests -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 1500 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -4 ${P}_pytest.log | cut -c1-220; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c16_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d.get("e2e"))
    print("decode", json.dumps(d.get("legs", {}).get("decode", d.get("decode")))[:1500])
except Exception as e:
    print("bench parse:", e)
PY
