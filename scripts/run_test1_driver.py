"""BASELINE config 1 (test/test1, 200x200, real(4) driver semantics) through the re-hosted reference driver on the GPU, timed
by the driver's own result.txt, next to the oracle's literal restatement on one host core at -O0 (what make-diagnosis.sh
builds) and -O3.  The reference's settings (r1 = r2 = 5e-3, max_iter 1e5, alpha 1) and, for the accelerated method, the same
settings.   python scripts/run_test1_driver.py  ->  gpurun_out/test1_driver.json"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
from tests.util import golden_json, ref_test1_inputs
from xlab_ee_fortran_b200 import _lib

exe = _lib.build_diagnose()
diag = golden_json()["reference_test1_diag_txt"]
A, B, C, bc = ref_test1_inputs()
out = {}


def run(env_extra, r8):
    with tempfile.TemporaryDirectory() as td:
        for n, arr in (("A.bin", A), ("B.bin", B), ("C.bin", C), ("bc_init.bin", bc)):
            arr.astype(np.float32).tofile(os.path.join(td, n))
        env = dict(os.environ, **env_extra)
        for k in range(2):      # second run: library and pool warm
            t = time.time()
            r = subprocess.run([exe] + (["--r8"] if r8 else []), input=diag, capture_output=True, text=True, cwd=td, env=env, timeout=600)
            wall = time.time() - t
        assert r.returncode == 0, r.stdout + r.stderr
        sec = float(open(os.path.join(td, "result.txt")).read().split(":")[1])
        sweeps = int([l for l in r.stdout.splitlines() if "Relaxation uses" in l][0].split()[2])
        return dict(result_txt_seconds=sec, process_wall_seconds=wall, sweeps=sweeps)


out["gpu_drop_in_reference_iteration_f32"] = run({}, False)
out["gpu_drop_in_reference_iteration_f64"] = run({}, True)
out["gpu_drop_in_line2_chebyshev_f64"] = run({"XEE_METHOD": "line2_chebyshev", "XEE_ARITH": "fast", "XEE_STALL_CHECKS": "20"}, True)
d = O.Domain((0.0, 1.0), (0.0, 1.0), 200, 200, 0, 0)
for variant in ("O0", "O3"):
    a, b, c = O.build_abc(A, B, C, d); g = O.geometry(d, np.float32)
    coe, _ = O.cal_coe(a, np.zeros_like(b), c, g["dr"], g["dz"], 200, 200, variant=variant)
    t = time.time()
    r = O.solve_elliptic(100000, 100, 10, 5, np.float32(5e-3), np.float32(5e-3), np.float32(1.0), bc, coe, -B, variant=variant)
    out[f"cpu_oracle_{variant}_one_core_f32"] = dict(seconds=time.time() - t, sweeps=int(r["max_iter"]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "test1_driver.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
