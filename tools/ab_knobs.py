"""A/B of the hand-off kernel's tuning knobs at 128 anchors (us per launch, MMVT and plain Langevin)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tools"))
os.environ["SWEEP_ANCHORS"] = "128"
import numpy as np, torch, bench
import seekr2_openmm_plugin_b200 as s2
from seekr2_openmm_plugin_b200 import synthetic as H
__file__ = os.path.join(%r, "tools", "decide_mode_sweep.py")
exec(open(__file__).read().split("out = {")[0])
print(json.dumps({"mmvt": us_per_step(128, "mmvt", "handoff"), "langevin": us_per_step(128, "langevin", "handoff")}))
''' % (ROOT, ROOT, ROOT)
for hs in ("1", "0"):
    for pf in ("1", "0"):
        env = dict(os.environ, SEEKR2_B200_HEAD_START=hs, SEEKR2_B200_PREFETCH=pf)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print("HEAD_START", hs, "PREFETCH", pf, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
