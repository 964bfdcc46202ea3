#!/bin/bash
# Quick GPU check: the fp4 convolution tests, then the default bench (JSON line into gpurun_out/$1_bench.json).
tag=${1:-quick}
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q -k "fp4" 2>&1 | tail -4
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["kernel_ms"].items()})
        print(d.get("parity"))
PY
