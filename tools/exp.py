#!/usr/bin/env python
"""Runs bench.py under several environment-variable settings and prints one line per variant (GPU box helper).
   python tools/exp.py "A=1" "A=1 B=2" ...   [--batch N]"""
import json, os, subprocess, sys
args = [a for a in sys.argv[1:] if not a.startswith("--")]
batch = "512"
for a in sys.argv[1:]:
    if a.startswith("--batch="):
        batch = a.split("=")[1]
for v in args or [""]:
    env = dict(os.environ)
    for kv in v.split():
        k, val = kv.split("=")
        env[k] = val
    r = subprocess.run([sys.executable, "bench.py", "--steps", "2", "--warmup", "2", "--batch", batch, "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        st = d["stages_ms_per_step"]
        print("%-40s value %8.1f  e2e %8.1f  ms/step %8.2f | %s" % (v or "(default)", d["value"], d["e2e"]["value"], d["ms_per_step"],
              " ".join("%s=%.1f" % (k[:6], x) for k, x in st.items())), flush=True)
    except Exception as e:
        print(v, "FAILED", e, r.stderr[-500:], flush=True)
