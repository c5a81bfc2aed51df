"""Second half of scripts/pin_against_reference.sh: compares output of the REAL reference (test/test1, real(4) build) with
tests/golden/golden.json (written by the oracle, tests/golden/make_golden.py).   pin_check.py <dir sweeps1000> <dir stop>
Exit 0 = every pin holds, 1 = mismatch."""
import hashlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["test1"]["f32"]
d1000, dstop = sys.argv[1], sys.argv[2]
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
rows, ok = [], True


def check(name, cond, detail):
    global ok
    ok &= bool(cond)
    rows.append(("ok  " if cond else "FAIL", name, detail))


def trace(path):
    """(iteration, err_now, ratio) triples of the debug-level-2 lines (elliptic_tools.f90:202-204)."""
    out = []
    for line in open(path, errors="replace"):
        m = re.search(r"Iter:\s*(\d+).*err_now:\s*([-+0-9.Ee]+).*ratio:\s*([-+0-9.Ee]+)", line)
        if m:
            out.append((int(m.group(1)), float(m.group(2)), float(m.group(3))))
    return out


# ---- fixed sweep count: bit-identical field expected (same arithmetic, same order: gfortran without -O is IEEE-exact)
psi = np.fromfile(os.path.join(d1000, "rchi-[BAROTROPIC]-O.bin"), np.float32)
check("rchi after 1000 sweeps: size", psi.size == 200 * 200, f"{psi.size}")
check("rchi after 1000 sweeps: sha256", sha(psi) == gold["psi1000_sha256"], f"{sha(psi)[:16]} vs golden {gold['psi1000_sha256'][:16]}")
l2 = float(np.sqrt((psi.astype(np.float64) ** 2).sum()))
check("rchi after 1000 sweeps: L2 (1e-6 rel)", abs(l2 - gold["psi1000_l2"]) <= 1e-6 * gold["psi1000_l2"], f"{l2:.9e} vs {gold['psi1000_l2']:.9e}")
tr = trace(os.path.join(d1000, "stdout.txt"))
gt = gold["trace_1000"]
check("residual trace of the first 1000 sweeps: 10 check lines", len(tr) == len(gt), f"{len(tr)} lines")
for (it, e, r), (git, ge, gr) in zip(tr, gt):
    # the reference prints ES12.3E2: three decimals
    check(f"  err_now at sweep {git}", it == git and abs(e - ge) <= 6e-4 * abs(ge), f"{e:.3e} vs {ge:.6e}")
# ---- to the stop rule (real(4): the stop sweep sits on the round-off floor, so sha first, then tolerances)
psi = np.fromfile(os.path.join(dstop, "rchi-[BAROTROPIC]-O.bin"), np.float32)
eta = np.fromfile(os.path.join(dstop, "eta-[BAROTROPIC]-A.bin"), np.float32)
same = sha(psi) == gold["psi_stop_sha256"]
check("rchi at the stop rule: sha256 (informational when the L2 test passes)", True, "identical" if same else "differs")
l2 = float(np.sqrt((psi.astype(np.float64) ** 2).sum()))
check("rchi at the stop rule: L2 (2e-4 rel)", abs(l2 - gold["psi_stop_l2"]) <= 2e-4 * gold["psi_stop_l2"], f"{l2:.9e} vs {gold['psi_stop_l2']:.9e}")
check("eta at the stop rule: max (2e-3 rel)", abs(float(eta.max()) - gold["eta_stop_max"]) <= 2e-3 * gold["eta_stop_max"], f"{float(eta.max()):.6e} vs {gold['eta_stop_max']:.6e}")
m = re.search(r"Relaxation uses\s+(\d+)", open(os.path.join(dstop, "stdout.txt"), errors="replace").read())
sweeps = int(m.group(1)) if m else -1
check("sweeps to the stop rule (informational: chaotic on the round-off floor)", True, f"{sweeps} vs golden {gold['stop_sweeps']}")
for r in rows:
    print(*r, sep="  ")
print("PINNED: the oracle reproduces the reference's output" if ok else "MISMATCH: the oracle (and everything tested against it) does not match the reference")
sys.exit(0 if ok else 1)
