"""The legacy TENDENCY efficiency decomposition (src/old-diagnose/diagnose.f90: nine elliptic solves per case,
efficiency.txt) through the re-hosted driver on the GPU against the oracle-side composition."""
import os
import subprocess

import numpy as np
import pytest

from tests.legacy_oracle import decompose
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def read_efficiency(path):
    """xtt-lib-python/XEffReader.py:3-32 restated (same prefixes, same split(':')/split(',') parsing)."""
    eff = dict(semi_internal=0.0, semi_cb1=0.0, internal=0.0, wtheta=0.0, local_response=0.0)
    for line in open(path):
        if line.startswith(' eta [L(B=0)    = 0]      w/  boundary'):
            eff['semi_internal'] += float((line.split(':')[1]).split(',')[1])
        elif line.startswith(' eta [L(B=0)    = dB]     wo/ boundary'):
            v = float((line.split(':')[1]).split(',')[1]); eff['semi_internal'] += v; eff['internal'] += v
        elif line.startswith(' eta [L(B=0)    = B0]     wo/ boundary'):
            v = float((line.split(':')[1]).split(',')[1]); eff['semi_internal'] += v; eff['internal'] += v
        elif line.startswith(' bndconv [L(B=0) = B0dB]   w/ boundary'):
            eff['semi_cb1'] += float((line.split(':')[1]).split(',')[1])
        elif line.startswith(' wtheta [L(B=0)    = J F] w/  boundary'):
            eff['wtheta'] += float((line.split(':')[1]).split(',')[1])
        elif line.startswith(' Local heat response (sum Q / sum dtheta_dt)'):
            eff['local_response'] += float(line.split(':')[1])
    eff['semi_total'] = eff['semi_internal'] + eff['semi_cb1']
    return eff


def test_legacy_tendency_decomposition(tmp_path):
    from xlab_ee_fortran_b200 import _lib
    from xlab_ee_fortran_b200 import workloads as W
    exe = _lib.build_old_diagnose()
    nr, nz = 56, 40
    Lr, Lz = (0.0, 5.0e5), (0.0, 1.4e4)
    A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
    r = np.linspace(*Lr, nr); z = np.linspace(*Lz, nz)
    rm = 0.5 * (r[:-1] + r[1:]); zm = 0.5 * (z[:-1] + z[1:])
    Rm, Zm = np.meshgrid(rm, zm)
    Q = (0.116 * np.exp(-((Rm - 4.0e4) / 3.0e4) ** 2 - ((Zm - 5.0e3) / 2.5e3) ** 2)).astype(np.float32)
    F = np.zeros_like(Q)      # pure heating: then the eta decomposition and the w*theta integral measure the same thing
    for name, arr in (("A.bin", A), ("B.bin", B), ("C.bin", C), ("Q.bin", Q), ("F.bin", F)):
        arr.astype(np.float32).tofile(tmp_path / name)
    dt_test = 3600.0
    cfg = ("CYLINDRICAL-TENDENCY-DENSITY_NORMAL-BARO_ALL // mode\n"
           f"{dt_test} // testing dt\n"
           f"{Lr[0]} {Lr[1]} {Lz[0]} {Lz[1]} // domain\n{nr} {nz} // grid\n. // in\n. // out\n"
           "A.bin // A\nB.bin // B\nC.bin // C\nQ.bin // Q\nF.bin // F\n"
           "1 1e-21 4000000 1.0 // rpsi: strategy residue max_iter alpha\n"
           "1 1e-18 4000000 1.0 // rchi\nno // rpsi bc\nno // rchi bc\n")
    rr = subprocess.run([exe, "--r8"], input=cfg, capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert rr.returncode == 0, rr.stdout + rr.stderr
    assert rr.stdout.count("Relaxation uses") == 7               # stage I + 4 chi solves + 2 integral-check solves
    ref = decompose(A, B, C, Q, F, Lr, Lz, dt_test, (1, 1e-21, 4000000, 1.0), (1, 1e-18, 4000000, 1.0), baro=2)
    txt = (tmp_path / "efficiency.txt").read_text()
    vals = {}
    for line in txt.splitlines():
        if ":" in line and not line.startswith(" #"):
            key, rest = line.split(":", 1)
            vals.setdefault(key.strip(), []).append([float(x) for x in rest.split(",")])
    assert vals["sum Q"][0][0] == pytest.approx(ref["sum_Q"], rel=1e-7)
    assert vals["sum dtheta_dt"][0][0] == pytest.approx(ref["sum_dtheta_dt"], rel=1e-6)
    pairs = {"eta [L(B=0)    = dB]     wo/ boundary": "sum_Qeta_0_dB", "eta [L(B=0)    = B0]     wo/ boundary": "sum_Qeta_0_B0",
             "eta [L(B=B0dB) = dB]     wo/ boundary": "sum_Qeta_B0dB_dB", "eta [L(B=B0dB) = B0]     wo/ boundary": "sum_Qeta_B0dB_B0",
             "wtheta [L(B=0)    = J F] w/  boundary": "sum_wtheta_0", "wtheta [L(B=B0dB) = J F] w/  boundary": "sum_wtheta_B0dB"}
    for label, key in pairs.items():
        got = vals[label][0]
        assert got[0] == pytest.approx(ref[key], rel=1e-6), label       # efficiency sums within 1e-6 relative
        assert got[1] == pytest.approx(ref[key] / ref["sum_Q"], rel=1e-6)
    # fields written by the driver (float32 files) against the oracle composition
    got = np.fromfile(tmp_path / "rpsi_before-O.bin", np.float32).reshape(nz, nr)
    assert rel_l2(got, ref["rpsi_before"]) < 1e-6
    assert rel_l2(np.fromfile(tmp_path / "theta_after-B.bin", np.float32).reshape(nz - 1, nr - 1), ref["theta_after"]) < 1e-6
    assert rel_l2(np.fromfile(tmp_path / "rchi-[B0dB_B0]-O.bin", np.float32).reshape(nz, nr), ref["rchi_B0dB_B0"]) < 1e-6
    # the reference's own reader parses the file
    eff = read_efficiency(tmp_path / "efficiency.txt")
    assert eff["internal"] == pytest.approx((ref["sum_Qeta_0_dB"] + ref["sum_Qeta_0_B0"]) / ref["sum_Q"], rel=1e-6)
    assert eff["wtheta"] == pytest.approx(ref["sum_wtheta_0"] / ref["sum_Q"], rel=1e-6)
    # the legacy driver's integral check: sum of the eta decomposition ~ w*theta integral (first-order agreement)
    assert eff["internal"] == pytest.approx(eff["wtheta"], rel=0.02)
    print("legacy decomposition: internal", eff["internal"], "wtheta", eff["wtheta"], "local response", eff["local_response"])


def test_legacy_boundary_conversion(tmp_path):
    """cal_exchange_conversion (old-diagnose/diagnose.f90:730-772, 1143-1174): with a rchi (and rpsi) boundary condition the
    legacy driver adds the top/bottom exchange term.  The re-hosted driver's bndconv lines of efficiency.txt and its
    bndconv*.bin files against the oracle-side routine (tests/legacy_oracle.exchange_conversion; real-typed r, dr, dz: [D5])."""
    from xlab_ee_fortran_b200 import _lib
    from xlab_ee_fortran_b200 import workloads as W
    exe = _lib.build_old_diagnose()
    nr, nz = 40, 28
    Lr, Lz = (0.0, 5.0e5), (0.0, 1.4e4)
    A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
    r = np.linspace(*Lr, nr); z = np.linspace(*Lz, nz)
    rm = 0.5 * (r[:-1] + r[1:]); zm = 0.5 * (z[:-1] + z[1:])
    Rm, Zm = np.meshgrid(rm, zm)
    Q = (0.116 * np.exp(-((Rm - 4.0e4) / 3.0e4) ** 2 - ((Zm - 5.0e3) / 2.5e3) ** 2)).astype(np.float32)
    F = np.zeros_like(Q)
    s = np.sin(np.pi * r / Lr[1])
    rchi_bc = np.zeros((nz, nr), np.float32); rchi_bc[-1] = 2.0e11 * s * (r / Lr[1]) ** 2        # chi on the top boundary
    rpsi_bc = np.zeros((nz, nr), np.float32); rpsi_bc[0] = 4.0e6 * s * (r / Lr[1]) ** 2          # pumping-like psi on the bottom
    for name, arr in (("A.bin", A), ("B.bin", B), ("C.bin", C), ("Q.bin", Q), ("F.bin", F), ("rchi_bc.bin", rchi_bc), ("rpsi_bc.bin", rpsi_bc)):
        arr.astype(np.float32).tofile(tmp_path / name)
    cfg = ("CYLINDRICAL-TENDENCY-DENSITY_NORMAL-BARO_ALL // mode\n3600.0 // testing dt\n"
           f"{Lr[0]} {Lr[1]} {Lz[0]} {Lz[1]} // domain\n{nr} {nz} // grid\n. // in\n. // out\n"
           "A.bin // A\nB.bin // B\nC.bin // C\nQ.bin // Q\nF.bin // F\n"
           "1 1e-21 4000000 1.0 // rpsi: strategy residue max_iter alpha\n"
           "1 1e-18 4000000 1.0 // rchi\nyes // rpsi bc\nrpsi_bc.bin\nyes // rchi bc\nrchi_bc.bin\n")
    rr = subprocess.run([exe, "--r8"], input=cfg, capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert rr.returncode == 0, rr.stdout + rr.stderr
    assert rr.stdout.count("Exchange conversion term check") == 2
    ref = decompose(A, B, C, Q, F, Lr, Lz, 3600.0, (1, 1e-21, 4000000, 1.0), (1, 1e-18, 4000000, 1.0), baro=2,
                    rchi_bc=rchi_bc, rpsi_bc=rpsi_bc)
    vals = {}
    for line in (tmp_path / "efficiency.txt").read_text().splitlines():
        if ":" in line and not line.startswith(" #"):
            key, rest = line.split(":", 1)
            vals.setdefault(key.strip(), []).append([float(x) for x in rest.split(",")])
    lines = {"bndconv [L(B=0) = B0dB]   w/ boundary": "sum_bndconv_0", "bndconv2 [L(B=0) = B0dB]   w/ boundary": "sum_bndconv2_0",
             "bndconv [L(B=B0dB) = B0dB]w/ boundary": "sum_bndconv_B0dB", "bndconv2 [L(B=B0dB) = B0dB]w/ boundary": "sum_bndconv2_B0dB"}
    for label, key in lines.items():
        got = vals[label][0]
        assert abs(ref[key]) > 0
        assert got[0] == pytest.approx(ref[key], rel=1e-5), label
        assert got[1] == pytest.approx(ref[key] / ref["sum_Q"], rel=1e-5), label
    for fn, key in (("bndconv-[0].bin", "bndconv_0"), ("bndconv2-[0].bin", "bndconv2_0"), ("bndconv-[B0dB].bin", "bndconv_B0dB"),
                    ("bndconv2-[B0dB].bin", "bndconv2_B0dB")):
        got = np.fromfile(tmp_path / fn, np.float32).reshape(2, nr - 1)
        assert np.abs(ref[key]).max() > 0 and rel_l2(got, ref[key]) < 1e-5, fn
    # the boundary-condition chi solves themselves, and the decomposition sum the driver prints (:806-808)
    assert rel_l2(np.fromfile(tmp_path / "rchi-[0_0]-O.bin", np.float32).reshape(nz, nr), ref["rchi_0_0"]) < 1e-6
    tot = ref["sum_Qeta_0_0"] + ref["sum_Qeta_0_dB"] + ref["sum_Qeta_0_B0"] + ref["sum_bndconv_0"]
    assert vals["etaQ [L(B=0)    = J F] w/  boundary"][0][0] == pytest.approx(tot, rel=1e-5)
