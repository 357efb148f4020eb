#!/bin/sh
# Round-end check on the GPU box: parity tests, smoke, both bench arms, every
# workload.  Lines land in gpurun_out/final_*.json.
if [ "$1" != notests ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
fi
t0=$(date +%s)
python bench.py > gpurun_out/final_C4.json 2> gpurun_out/final_C4.err
t1=$(date +%s)
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
t2=$(date +%s)
echo "default bench took $((t1 - t0)) s, reference arm $((t2 - t1)) s"
tail -c 600 gpurun_out/final_ref.json
if [ "$1" != notests ]; then
for w in C1 C2 C3 C5; do python bench.py --workload $w > gpurun_out/final_$w.json 2>/dev/null; done
fi
python - <<PY
import json
for w in ("C4", "C2", "C5", "C1", "C3"):
    d = json.loads(open("gpurun_out/final_%s.json" % w).readlines()[-1])
    print(w, round(d["value"], 2), round(d["ms_per_step"], 3),
          round(d["e2e"]["value"], 2),
          round(d.get("e2e_packed8", {}).get("value", 0), 2),
          d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
          round(d["chain_roofline"]["frac"], 3), d["gpu_launches"],
          d["cpu_baseline"]["value"] if d["cpu_baseline"] else None,
          d["clocks"])
PY
