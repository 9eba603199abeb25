"""Runs every GPU parity check without stopping at the first failure and writes gpurun_out/gpu_checks.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gpu_checks  # noqa: E402
import gpu_checks_train  # noqa: E402,F401
import gpu_checks_strict  # noqa: E402,F401
import gpu_checks_data  # noqa: E402,F401

if __name__ == "__main__":
    only = sys.argv[1:]
    if only:
        gpu_checks.CHECKS[:] = [c for c in gpu_checks.CHECKS if any(o in c.__name__ for o in only)]
    res = gpu_checks.run_all()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_checks.json"), "w") as f:
        json.dump(res, f, indent=1)
    bad = [r for r in res if not r["ok"]]
    print(f"{len(res) - len(bad)} passed, {len(bad)} failed")
    sys.exit(1 if bad else 0)
