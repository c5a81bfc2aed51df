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
