"""Generates the committed golden vectors in tests/golden/ (run from the repo root, in the
build container where /root/reference is mounted; nothing at test time reads /root/reference).

The reference commits no expected outputs and no Fortran compiler exists
here, so golden vectors are produced by the independent numpy restatement oracle/numpy_ref.py;
the C++ oracle (oracle/xee_oracle.hpp) must reproduce them bit for bit (tests/test_oracle.py),
and the CUDA path is then checked against the C++ oracle.  The sha256 of the reference's own
test1 input files is recorded so tests can prove their regenerated inputs are byte-identical.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import numpy_ref as N  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_T1 = "/root/reference/test/test1"


def test1_inputs():
    """test/test1/test-setup.py:20-55 restated."""
    nr = nz = 200
    r = np.linspace(0.0, 1.0, nr); z = np.linspace(0.0, 1.0, nz)
    A = np.ones((nz, nr), np.float32); C = np.ones((nz, nr), np.float32)
    bc = np.zeros((nz, nr), np.float32)
    rr, zz = np.meshgrid(r, z)
    B = (1e-2 * np.sin(2.0 * np.pi * (rr - r[0]) / 1.0) * np.sin(3.0 * np.pi * (zz - z[0]) / 1.0)).astype(np.float32)
    return A, B, C, bc


def small_case(dt):
    """A small baroclinic case with non-zero Dirichlet data, B != 0 and alpha < 1."""
    nr, nz = 40, 32
    rng = np.random.default_rng(20140911)
    Lr, Lz = (0.0, 2.0e5), (0.0, 1.2e4)
    r = np.linspace(*Lr, nr); z = np.linspace(*Lz, nz)
    rr, zz = np.meshgrid(r, z)
    A = (1.0e-4 * (1.0 + 0.3 * np.cos(zz / Lz[1] * np.pi))).astype(np.float32)
    C = (1.0e-8 * (1.0 + 4.0 * np.exp(-(rr / 5.0e4) ** 2))).astype(np.float32)
    B = (2.0e-7 * np.sin(np.pi * rr / Lr[1]) * np.sin(2 * np.pi * zz / Lz[1])).astype(np.float32)
    f = (1e-9 * rng.standard_normal((nz, nr))).astype(np.float32)
    bc = (1e-2 * rng.standard_normal((nz, nr))).astype(np.float32)
    g = N.geometry(Lr, Lz, nr, nz, dt)
    return dict(nr=nr, nz=nz, Lr=Lr, Lz=Lz, A=A, B=B, C=C, f=f, bc=bc, g=g)


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    meta = {}
    A, B, C, bc = test1_inputs()
    if os.path.isdir(REF_T1):
        meta["reference_test1_sha256"] = {n: sha(os.path.join(REF_T1, n)) for n in ("A.bin", "B.bin", "C.bin", "bc_init.bin")}
        for n, arr in (("A.bin", A), ("B.bin", B), ("C.bin", C), ("bc_init.bin", bc)):
            assert hashlib.sha256(arr.tobytes()).hexdigest() == meta["reference_test1_sha256"][n], n
        meta["reference_test1_diag_txt"] = open(os.path.join(REF_T1, "diag.txt"), newline="").read()
    # ---- test1 (BASELINE config 1): BAROTROPIC, f = -B, bc = 0, alpha = 1, r1 = r2 = 5e-3
    t1 = {}
    for name, dt, full in (("f32", np.float32, True), ("f64", np.float64, True)):
        g = N.geometry((0.0, 1.0), (0.0, 1.0), 200, 200, dt)
        a, b, c = N.build_abc(A.astype(dt), B.astype(dt), C.astype(dt), g)
        coe = N.cal_coe(a, np.zeros_like(b), c, g["dr"], g["dz"])
        f = -(B.astype(dt))
        r = N.solve_elliptic(1000, 100, 10, 5, 5e-3, 5e-3, 1.0, bc.astype(dt), coe, f, snapshots=(100, 1000))
        psi = r["dat"]
        ent = dict(
            trace_1000=r["trace"],
            psi1000_min=float(psi.min()), psi1000_max=float(psi.max()),
            psi1000_l2=float(np.sqrt((psi.astype(np.float64) ** 2).sum())),
            psi1000_100_100=float(psi[99, 99]),
            psi1000_sha256=hashlib.sha256(psi.tobytes()).hexdigest(),
            psi100_sha256=hashlib.sha256(r["snaps"][100].tobytes()).hexdigest(),
            coe_sha256=hashlib.sha256(coe.tobytes()).hexdigest(),
        )
        if full:
            rf = N.solve_elliptic(100000, 100, 10, 5, 5e-3, 5e-3, 1.0, bc.astype(dt), coe, f)
            pf = rf["dat"]
            eta = N.cal_eta(pf, g)
            ent.update(stop_sweeps=rf["max_iter"], stop_r1=rf["r1"], stop_r2=rf["r2"], stop_err=rf["err"],
                       psi_stop_min=float(pf.min()), psi_stop_max=float(pf.max()),
                       psi_stop_l2=float(np.sqrt((pf.astype(np.float64) ** 2).sum())),
                       psi_stop_sha256=hashlib.sha256(pf.tobytes()).hexdigest(),
                       eta_stop_min=float(eta.min()), eta_stop_max=float(eta.max()),
                       eta_stop_sha256=hashlib.sha256(eta.tobytes()).hexdigest(),
                       trace_stop_tail=rf["trace"][-12:])
            np.save(os.path.join(OUT, f"test1_psi_stop_{name}_ds.npy"), pf[::8, ::8].copy())
        t1[name] = ent
        print(name, {k: v for k, v in ent.items() if not k.startswith("trace")})
    meta["test1"] = t1
    # ---- small baroclinic case, full fields
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        sc = small_case(dt)
        g = sc["g"]
        a, b, c = N.build_abc(sc["A"].astype(dt), sc["B"].astype(dt), sc["C"].astype(dt), g)
        coe = N.cal_coe(a, b, c, g["dr"], g["dz"])
        f = sc["f"].astype(dt)
        Lpsi = N.do_elliptic(sc["bc"].astype(dt), coe)
        r50 = N.solve_elliptic(50, 10, 3, 2, 1e-30, 1.0, 0.8, sc["bc"].astype(dt), coe, f)
        # r1 alone decides: r2 >= 1 never blocks
        rms_f = float(np.sqrt((f[1:-1, 1:-1].astype(np.float64) ** 2).mean()))
        rstop = N.solve_elliptic(200000, 50, 4, 3, 1e-3 * rms_f, 2.0, 0.8, sc["bc"].astype(dt), coe, f)
        eta = N.cal_eta(rstop["dat"], g)
        u, w = N.cal_uw(rstop["dat"], g)
        np.savez_compressed(
            os.path.join(OUT, f"small_{name}.npz"),
            A=sc["A"], B=sc["B"], C=sc["C"], f=sc["f"], bc=sc["bc"], Lr=np.array(sc["Lr"]), Lz=np.array(sc["Lz"]),
            ra=g["ra"], za=g["za"], rho=g["rho"], exner=g["exner"], dr=g["dr"], dz=g["dz"],
            a=a, b=b, c=c, coe=coe, Lpsi=Lpsi, psi50=r50["dat"], trace50=np.array(r50["trace"]),
            psi_stop=rstop["dat"], stop_sweeps=rstop["max_iter"], stop_r1=rstop["r1"], stop_r2=rstop["r2"],
            stop_err=rstop["err"], trace_stop=np.array(rstop["trace"]), eta=eta, u=u, w=w,
            r1_in=1e-3 * rms_f)
        print("small", name, "stop", rstop["max_iter"], rstop["r1"], rstop["err"])
    json.dump(meta, open(os.path.join(OUT, "golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
