#!/bin/bash
# One GPU session of the development loop: the parity suite, the C2 line with its block-mix sweep, C3 / C4 per coder and mixed.
# usage (on the GPU box, from the repo root): tools/gpu_round.sh <tag>
T=${1:-x}; O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
python bench.py --no-codecs --no-cpu-baseline --steps 10 2>/dev/null | grep "^{" > $O/${T}_c2.json
for m in 100,0,0 0,100,0 0,0,100; do
  python bench.py --no-codecs --no-cpu-baseline --e2e-steps 0 --steps 10 --c2-mix $m 2>/dev/null | grep "^{" > $O/${T}_c2_sweep_$m.json
done
for w in c3 c4; do for v in 2 4 2,4; do
  python bench.py --workload $w --steps 3 --e2e-steps 0 --no-cpu-baseline --sp-versions $v 2>/dev/null | grep "^{" > $O/${T}_${w}_v$v.json
done; done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_c*.json")):
    try:
        d = json.loads(open(f).read())
        r = d["roofline"]
        print(f.split("/")[-1], "value %.0f ms %.3f" % (d["value"], d["ms_per_step"]), "frac %.3f" % r["frac"] if r.get("frac") else "cyc/sym %.0f" % r.get("cycles_per_symbol", 0), "e2e", d["e2e"].get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
