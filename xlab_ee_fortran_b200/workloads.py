"""Synthetic inputs for the BASELINE configs: a balanced vortex (A, B, C fields in
the reference's float32 file format) and heating-location lattices.

The vortex follows xtt-lib-python/XWindProfile.py:10-23 (piecewise-constant absolute vorticity:
M_k^2 = f_k^2 r^4/4 + K_k, continuous across the ring radii) multiplied by a vertical decay
D(z) = exp(-z/H).  A = N^2, C = r^-3 d(M^2)/dr (inertial stability), B = -r^-3 d(M^2)/dz (thermal
wind), clipped so that A C - B^2 >= 0.05 A C (the operator stays elliptic).
Host-side input generation only; no solver arithmetic lives here.
"""
from __future__ import annotations

import numpy as np


class WindProfile:
    """xtt-lib-python/XWindProfile.py:1-23 restated (vectorised over r)."""

    def __init__(self, f0, f_arr, radius_arr):
        self.f0 = float(f0); self.f_arr = list(f_arr); self.radius_arr = list(radius_arr)
        self.konst = [0.0] * len(self.f_arr)
        for i in range(1, len(self.konst)):       # XWindProfile.py:12-13
            self.konst[i] = self.konst[i - 1] + (self.radius_arr[i - 1] ** 4.0) / 4.0 * (self.f_arr[i - 1] ** 2.0 - self.f_arr[i] ** 2.0)

    def region(self, r):
        reg = np.full(np.shape(r), len(self.f_arr) - 1, int)
        for i in range(len(self.radius_arr) - 1, -1, -1):
            reg = np.where(np.asarray(r) < self.radius_arr[i], i, reg)
        return reg

    def m_base_sq(self, r):
        """M_k^2 = f_k^2 r^4/4 + K_k of the barotropic profile (absolute angular momentum squared)."""
        reg = self.region(r)
        fk = np.asarray(self.f_arr)[reg]; kk = np.asarray(self.konst)[reg]
        return fk ** 2 * np.asarray(r) ** 4 / 4.0 + kk, fk

    def getWind(self, r):                           # XWindProfile.py:16-23
        r = np.asarray(r, float)
        m2, _ = self.m_base_sq(r)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.sqrt(m2) / r - 0.5 * self.f0 * r
        return np.where(r != 0.0, v, 0.0)


def vortex_fields(nr, nz, Lr=(0.0, 1.0e6), Lz=(0.0, 1.5e4), f0=5e-5, f_arr=(1.0e-3, 5e-5), radius_arr=(5.0e4,),
                  H=1.0e4, N2=1.0e-4, thermal_wind_consistent=True):
    """Returns float32 (nz, nr) arrays A, B, C as the reference's A.bin/B.bin/C.bin would hold them.

    thermal_wind_consistent: A = N2 - int_0^r dB/dz dr' so that A = (g/theta0) dtheta/dz and
    B = -(g/theta0) dtheta/dr derive from ONE theta field; only then does the legacy driver's
    "Integral check" identity  int(Q eta) = (g/theta0) int(w theta)  hold (old-diagnose/diagnose.f90:677-725).
    """
    r = np.linspace(Lr[0], Lr[1], nr); z = np.linspace(Lz[0], Lz[1], nz)
    wp = WindProfile(f0, f_arr, radius_arr)
    m2b, fk = wp.m_base_sq(r)
    mb = np.sqrt(m2b)
    m = mb - 0.5 * f0 * r ** 2                      # r * v(r)
    with np.errstate(divide="ignore", invalid="ignore"):
        dm = np.where(mb > 0, 0.5 * fk ** 2 * r ** 3 / mb, 0.0) - f0 * r      # d(r v)/dr
    D = np.exp(-z / H)[:, None]; dD = -D / H
    M = m[None, :] * D + 0.5 * f0 * r[None, :] ** 2
    dM_dr = dm[None, :] * D + f0 * r[None, :]
    dM_dz = m[None, :] * dD
    F = f0 + (fk[None, :] - f0) * D                                      # r -> 0 limit of the local vorticity
    with np.errstate(divide="ignore", invalid="ignore"):
        C = np.where(r[None, :] > 0, 2.0 * M * dM_dr / r[None, :] ** 3, F ** 2)
        B = np.where(r[None, :] > 0, -2.0 * M * dM_dz / r[None, :] ** 3, 0.0)
    C = np.maximum(C, f0 ** 2)
    A = np.full((nz, nr), N2)
    lim = np.sqrt(0.95 * A * C)                                         # A C - B^2 >= 0.05 A C
    B = np.clip(B, -lim, lim)
    if thermal_wind_consistent:
        dBdz = np.gradient(B, z, axis=0)
        A = N2 - np.concatenate([np.zeros((nz, 1)), np.cumsum(0.5 * (dBdz[:, 1:] + dBdz[:, :-1]) * np.diff(r)[None, :], axis=1)], axis=1)
        A = np.maximum(A, 0.5 * N2)
    return A.astype(np.float32), B.astype(np.float32), C.astype(np.float32)


def heating_lattice(n_r, n_z, Lr, Lz, sigma_r, sigma_z, q0=None, r_frac=(0.0, 1.0), z_frac=(0.0, 1.0)):
    """[n_r*n_z, 5] rows {r_c, z_c, sigma_r, sigma_z, Q0}: centres (k+0.5)/n of the (sub)domain; Q0 = Cp * 10 K/day."""
    if q0 is None:
        q0 = 3.5 * 287.0 * 10.0 / 86400.0
    r0, r1 = Lr[0] + r_frac[0] * (Lr[1] - Lr[0]), Lr[0] + r_frac[1] * (Lr[1] - Lr[0])
    z0, z1 = Lz[0] + z_frac[0] * (Lz[1] - Lz[0]), Lz[0] + z_frac[1] * (Lz[1] - Lz[0])
    rc = r0 + (np.arange(n_r) + 0.5) * (r1 - r0) / n_r
    zc = z0 + (np.arange(n_z) + 0.5) * (z1 - z0) / n_z
    R, Z = np.meshgrid(rc, zc, indexing="ij")
    out = np.zeros((n_r * n_z, 5))
    out[:, 0] = R.ravel(); out[:, 1] = Z.ravel(); out[:, 2] = sigma_r; out[:, 3] = sigma_z; out[:, 4] = q0
    return out


def test1_style_fields(nr, nz):
    """The reference's test/test1 input formulas at an arbitrary grid (unit square, A=C=1; test-setup.py:42-55)."""
    r = np.linspace(0.0, 1.0, nr); z = np.linspace(0.0, 1.0, nz)
    rr, zz = np.meshgrid(r, z)
    A = np.ones((nz, nr), np.float32); C = np.ones((nz, nr), np.float32)
    B = (1e-2 * np.sin(2.0 * np.pi * rr) * np.sin(3.0 * np.pi * zz)).astype(np.float32)
    return A, B, C


class Pumping:
    """xtt-lib-python/XPumping.py:32-103 restated: piecewise-quadratic Ekman pumping rho*w(r) on [r0,r1,r2] and its
    integral r*psi(r), used as the bottom boundary condition of r*psi (BASELINE config 5)."""

    def __init__(self, rho_w0, r_arr):
        if len(r_arr) != 3:
            raise Exception("The length of r array must be exactly 3, the input length is %d" % (len(r_arr),))
        self.rho_w0 = float(rho_w0); self.r_arr = np.array(r_arr, float)
        r0, r1, r2 = self.r_arr
        self.coe = np.zeros((2, 2))
        self.coe[0][0] = -4.0 * self.rho_w0 / (r1 - r0) ** 2.0                     # XPumping.py:61
        self.coe[0][1] = -self.coe[0][0] * self.int_part(r0, r0, r1)               # :62
        a = np.array([[self.int_part(r2, r1, r2), 1.0], [self.int_part(r1, r1, r2), 1.0]])
        b = np.array([0.0, self.coe[0][0] * self.int_part(r1, r0, r1) + self.coe[0][1]])
        self.coe[1][0], self.coe[1][1] = np.linalg.solve(a, b)                     # :65-76

    @staticmethod
    def int_part(at_r, r_min, r_max):                                              # :40-41
        return (at_r ** 4.0) / 4.0 - (r_min + r_max) / 3.0 * (at_r ** 3.0) + r_min * r_max * (at_r ** 2.0) / 2.0

    def getRPsi(self, r):                                                          # :79-90
        r = np.asarray(r, float); r0, r1, r2 = self.r_arr
        inner = self.coe[0][0] * self.int_part(r, r0, r1) + self.coe[0][1]
        outer = self.coe[1][0] * self.int_part(r, r1, r2) + self.coe[1][1]
        return np.where(r <= r0, 0.0, np.where(r <= r1, inner, np.where(r <= r2, outer, 0.0)))

    def getRhoW(self, r):                                                          # :92-103
        r = np.asarray(r, float); r0, r1, r2 = self.r_arr
        inner = self.coe[0][0] * (r - r0) * (r - r1); outer = self.coe[1][0] * (r - r1) * (r - r2)
        return np.where(r <= r0, 0.0, np.where(r <= r1, inner, np.where(r <= r2, outer, 0.0)))

    def getTotalFlux(self):                                                        # :48-49
        r0, r1, _ = self.r_arr
        return self.coe[0][0] * (self.int_part(r1, r0, r1) - self.int_part(r0, r0, r1))


# ---- BASELINE config 5: time series of synthetic vortex snapshots
SERIES_COLS = ("f0", "f_core", "f_env", "radius", "konst1", "H", "N2", "pump_r0", "pump_r1", "pump_r2", "pump_c00",
               "pump_c01", "pump_c10", "pump_c11", "heat_rc", "heat_zc", "heat_sr", "heat_sz", "heat_q0", "fric_k", "fric_h")


def series_params(n_snap, total=None, first=0):
    """[n_snap, 21] rows (SERIES_COLS) for snapshots first..first+n_snap-1 of a `total`-long series:
    core vorticity 1e-3*(0.5+s), ring radius 5e4*(1.5-0.5 s), pumping 0.01*(1+s), s = n/(total-1);
    heating blob and friction F = -k v(r) exp(-z/h) fixed."""
    total = total or n_snap
    out = np.zeros((n_snap, len(SERIES_COLS)))
    for q in range(n_snap):
        s = (first + q) / max(total - 1, 1)
        f0, f_env = 5e-5, 5e-5
        f_core = 1.0e-3 * (0.5 + s); radius = 5.0e4 * (1.5 - 0.5 * s)
        wp = WindProfile(f0, [f_core, f_env], [radius])
        pm = Pumping(0.01 * (1.0 + s), [0.0, 5e4, 2e5])
        out[q] = [f0, f_core, f_env, radius, wp.konst[1], 1.0e4, 1.0e-4, *pm.r_arr, pm.coe[0][0], pm.coe[0][1], pm.coe[1][0],
                  pm.coe[1][1], 4.0e4, 5.0e3, 1.0e4, 2.0e3, 3.5 * 287.0 * 10.0 / 86400.0, 1e-5, 1.0e3]
    return out


def series_fields_host(row, nr, nz, Lr, Lz):
    """Host (numpy) statement of what the device builders compute for one snapshot row:
    A, B, C (float32, as files), bottom boundary r*psi(r, z=Lz0) (float64, O grid row 0) and F on B (float64)."""
    p = dict(zip(SERIES_COLS, row))
    A, B, C = vortex_fields(nr, nz, Lr, Lz, f0=p["f0"], f_arr=(p["f_core"], p["f_env"]), radius_arr=(p["radius"],), H=p["H"],
                            N2=p["N2"], thermal_wind_consistent=False)
    r = np.linspace(Lr[0], Lr[1], nr); z = np.linspace(Lz[0], Lz[1], nz)
    pm = Pumping.__new__(Pumping)
    pm.r_arr = np.array([p["pump_r0"], p["pump_r1"], p["pump_r2"]]); pm.coe = np.array([[p["pump_c00"], p["pump_c01"]], [p["pump_c10"], p["pump_c11"]]])
    bottom = pm.getRPsi(r)
    wp = WindProfile(p["f0"], [p["f_core"], p["f_env"]], [p["radius"]])
    rm = 0.5 * (r[:-1] + r[1:]); zm = 0.5 * (z[:-1] + z[1:])
    F = -p["fric_k"] * wp.getWind(rm)[None, :] * np.exp(-zm / p["fric_h"])[:, None]
    return A, B, C, bottom, F
