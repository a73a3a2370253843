import json
import re
import sys

for line in sys.stdin:
    m = re.search(r"\{.*\}", line)
    if not m:
        print(line.strip())
        continue
    d = json.loads(m.group(0))
    print({k: d.get(k) for k in ("sweeps", "adjoint_sweeps", "same_solver_path", "bit_identical", "max_rel", "J_rel", "verified", "host_path_equal", "ok")})
