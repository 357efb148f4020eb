#!/bin/sh
# Round-end check on the GPU box: parity tests, smoke, both bench arms.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/final_C2.json 2> gpurun_out/final_C2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
tail -c 500 gpurun_out/final_ref.json
for w in C4 C5; do python bench.py --workload $w > gpurun_out/final_$w.json 2>/dev/null; done
python - <<PY
import json
for w in ("C2", "C4", "C5"):
    d = json.loads(open("gpurun_out/final_%s.json" % w).readlines()[-1])
    print(w, round(d["value"], 2), round(d["ms_per_step"], 3),
          round(d["e2e"]["value"], 2), round(d["e2e_packed8"]["value"], 2),
          round(d["roofline"]["frac"], 3), round(d["chain_roofline"]["frac"], 3),
          d["gpu_launches"],
          d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["clocks"])
PY
