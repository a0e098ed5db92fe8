"""Generates tests/golden/ref_cuda_platform/: outputs of the REFERENCE'S OWN CUDA platform (CudaSeekr2Kernels.cpp + its
kernels through NVRTC, unmodified, compiled against tests/openmm_shim by oracle/build_shim.sh) for the cases of
tests/ref_platform_cases.py.  Needs a GPU and oracle/_ref/shim/ref_scenario:

    gpurun -- 'python tests/golden/make_ref_platform_golden.py gpurun_out/ref_golden'   # then copy into tests/golden/ref_cuda_platform/

Per case: <name>.csv (crossings file, bytes as the reference wrote them), <name>.stats (statistics file, MMVT), and in
manifest.json the sha256 of the final posq | posqCorrection | velm bytes of the real atoms, the context time and step
count.  Inputs are rebuilt from seeds (tests/helpers.py), so only outputs are stored."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402
import ref_platform_cases as RC  # noqa: E402
from shim_io import read_dump, write_scenario  # noqa: E402


def state_digest(hb, n):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(hb.posq[:n]).tobytes())
    if hb.corr is not None:
        h.update(np.ascontiguousarray(hb.corr[:n]).tobytes())
    h.update(np.ascontiguousarray(hb.velm[:n]).tobytes())
    return h.hexdigest()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    binary = os.path.join(H.ROOT, "oracle", "_ref", "shim", "ref_scenario")
    manifest = {}
    for name, case in sorted(RC.cases().items()):
        syn = case["make"]()
        n = syn.positions.shape[0]
        with tempfile.TemporaryDirectory() as tmp:
            scen, cross, stats, dump = (os.path.join(tmp, f) for f in ("scenario.txt", "crossings.txt", "stats.txt", "dump.bin"))
            write_scenario(scen, syn, case["kind"], case["precision"], case["steps"], RC.pool_for(case, syn), output=cross, dump=dump,
                           savestats=stats if case.get("stats") else "", counter=case["counter"], end_on_src=case.get("end_on_src"),
                           milestones=RC.labels_for(case, syn) if case["kind"].startswith("mmvt") else (),
                           srcs=case.get("srcs", ()), dests=case.get("dests", ()))
            proc = subprocess.run([binary, scen], capture_output=True, text=True, timeout=900)
            if proc.returncode != 0:
                raise SystemExit(f"{name}: ref_scenario failed\n{proc.stdout}\n{proc.stderr}")
            hb, time, step_count = read_dump(dump, case["precision"], n)
            with open(os.path.join(out_dir, name + ".csv"), "wb") as f:
                f.write(open(cross, "rb").read())
            entry = {"state_sha256": state_digest(hb, n), "time": time, "step_count": step_count,
                     "crossing_lines": sum(1 for ln in open(cross) if not ln.startswith("#"))}
            if case.get("stats") and os.path.exists(stats):
                with open(os.path.join(out_dir, name + ".stats"), "wb") as f:
                    f.write(open(stats, "rb").read())
                entry["has_stats"] = True
            manifest[name] = entry
            print(name, entry)
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_cuda_platform"))
