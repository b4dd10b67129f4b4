#!/usr/bin/env python
"""Golden vectors from the one runnable piece of the reference that reads this path's file format:
bin/pqrs_center.py (pure Python).  Runs ONLY in the build container (needs /root/reference); it is executed
unmodified, as a subprocess, on the committed .pqrs fixtures, and its stdout (header line + '%g %g %g' centre)
is committed as tests/golden/pqrs_center.json.  tests/test_molfile.py then demands that the library's own
readers see the same header and the same centre: the reference's tool accepts our files and agrees on them."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
out = {}
for name in ("docked", "ligdecs", "minimized", "xtal_rec"):
    r = subprocess.run([sys.executable, "/root/reference/bin/pqrs_center.py", os.path.join(GOLD, name + ".pqrs")],
                       capture_output=True, text=True, check=True)
    out[name] = r.stdout.strip().split("\n")
json.dump(out, open(os.path.join(GOLD, "pqrs_center.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
