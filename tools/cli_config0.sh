#!/bin/bash
# End-to-end command line on BASELINE.json configs[0]: 5 000 random 3-SAT problems (n = 100, m/n = 4.2) as a compact-JSON
# file -> satyr.py (T = 1 000, W = 100) -> JSON output.  Prints the predict time the CLI logs and the solved count.
set -e
f=${TMPDIR:-/tmp}/config0_5000.json
python - "$f" <<'PY'
import json, sys, numpy as np
rng = np.random.Generator(np.random.PCG64(1000))
with open(sys.argv[1], "w") as out:
    for j in range(5000):
        n, m = 100, 420
        v = np.stack([rng.choice(n, size=3, replace=False) + 1 for _ in range(m)])
        s = rng.integers(0, 2, size=(m, 3)) * 2 - 1
        out.write(json.dumps([[n, m], (v * s).reshape(-1).tolist(), np.repeat(np.arange(1, m + 1), 3).tolist(), 1.0, ["p%d" % j]]) + "\n")
PY
for rep in 1 2; do
  python satyr.py config/Predict/sp.yaml "$f" 1000 -w 100 -e 0.5 -s 1 -v -o "$f.out" 2>&1 | grep -i "time spent" || true
done
python - "$f.out" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
print("problems %d, solved %d" % (len(rows), sum(r["solved"] for r in rows)))
PY
