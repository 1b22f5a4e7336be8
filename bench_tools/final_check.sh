#!/bin/bash
# GPU-side regression of a round: parity suite, smoke, one full bench line (run under gpurun)
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_final.log").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["roofline"]["frac"],
      l["roofline"]["achieved"], l["launches_per_step"], l["clocks"])
print(l["cpu_baseline"]["value"]); print(l["kpconv_stack"]); print(l["c4"])
print(l["kpconv"]["fwd_ms"], l["kpconv"]["fwd_bwd_ms"])
for k, v in l["operators"].items():
    print(k, v)
PY
