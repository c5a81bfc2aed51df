"""Time-series diagnosis (Part 4 of include/xee_b200.h, BASELINE config 5): one operator per snapshot."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .efficiency_map import _prm
from .plan import ARITH_FAST, ARITH_STRICT, CHEBYSHEV, F32, F64, JACOBI, METHODS, SolveParams

COLS = ("iters", "r1", "err", "sum_Q", "ke_gen", "efficiency", "w_absmax", "u_absmax")


class _SeriesDesc(C.Structure):
    _fields_ = [("dtype", C.c_int), ("nr", C.c_int), ("nz", C.c_int), ("nsnap", C.c_int), ("density_mode", C.c_int),
                ("arith", C.c_int), ("method", C.c_int), ("device", C.c_int), ("Lr", C.c_double * 2), ("Lz", C.c_double * 2),
                ("r1_rel_rms_f", C.c_double)]


class TimeSeries:
    def __init__(self, nr, nz, Lr, Lz, nsnap, dtype="f64", density_mode=0, arith="fast", method="line_chebyshev", r1_rel=1e-12,
                 device=-1):
        _lib.require_gpu()
        self.nr, self.nz, self.nsnap = int(nr), int(nz), int(nsnap)
        self.np_dtype = np.float64 if dtype == "f64" else np.float32
        d = _SeriesDesc(F64 if dtype == "f64" else F32, nr, nz, nsnap, density_mode,
                        ARITH_STRICT if arith == "strict" else ARITH_FAST, METHODS[method],
                        device, (C.c_double * 2)(*Lr), (C.c_double * 2)(*Lz), float(r1_rel))
        self._h = C.c_void_p()
        _lib.check(_lib.lib().xee_series_create(C.byref(d), C.byref(self._h)), "series_create")

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.lib().xee_series_destroy(self._h)
            self._h = None

    __del__ = close

    def run(self, params, p: SolveParams):
        params = np.ascontiguousarray(params, np.float64)
        assert params.shape == (self.nsnap, 21)
        table = np.zeros((self.nsnap, len(COLS)))
        q = _prm(p)
        _lib.check(_lib.lib().xee_series_run_host(self._h, params.ctypes.data_as(C.c_void_p), C.byref(q),
                                                  table.ctypes.data_as(C.c_void_p)), "series_run_host")
        return table

    def field(self, which):
        nr, nz, n = self.nr, self.nz, self.nsnap
        idx, shape = {"psi": (0, (n, nz, nr)), "f": (1, (n, nz, nr)), "theta": (2, (n, nz - 1, nr - 1)), "u": (3, (n, nz - 1, nr)),
                      "w": (4, (n, nz, nr - 1)), "A": (5, (n, nz, nr)), "B": (6, (n, nz, nr)), "C": (7, (n, nz, nr)),
                      "m2": (8, (n, nz - 1, nr - 1))}[which]
        out = np.zeros(shape, self.np_dtype)
        _lib.check(_lib.lib().xee_series_get_field(self._h, idx, out.ctypes.data_as(C.c_void_p)), "series_get_field")
        return out

    def sweep_kernel_stats(self, reset=False):
        ms = C.c_double(0); n = C.c_longlong(0)
        _lib.lib().xee_series_sweep_kernel_stats(self._h, C.byref(ms), C.byref(n), C.c_int(int(reset)))
        if reset:
            _lib.lib().xee_series_probe_stats(self._h, None, C.c_int(1))
        return ms.value, n.value

    def probe_ms(self):
        """Host wall time (ms) of the spectral-radius probes since the last stats reset."""
        ms = C.c_double(0)
        _lib.lib().xee_series_probe_stats(self._h, C.byref(ms), C.c_int(0))
        return ms.value

    def kernel_info(self):
        """(variant 1..5, sweeps per kernel launch, kernel launches since the last stats reset) of the sweep kernel."""
        v = C.c_int(0); d = C.c_int(0); n = C.c_longlong(0)
        _lib.lib().xee_series_kernel_info(self._h, C.byref(v), C.byref(d), C.byref(n))
        return v.value, d.value, n.value


def run_sharded(nr, nz, Lr, Lz, total, p: SolveParams, world=1, rank=0, device=-1, **kw):
    """BASELINE config 5 on `world` GPUs: snapshots [a, b) of a `total`-long series on this rank (static contiguous chunks,
    SURVEY section 8e: independent solves, nothing exchanged), parameters generated per rank from the snapshot index.
    Returns (a, b, table [b-a, 8]); gather the rows with efficiency_map.gather_rows (one collective)."""
    from . import workloads as W
    from .efficiency_map import partition
    a, b = partition(total, world, rank)
    ts = TimeSeries(nr, nz, Lr, Lz, b - a, device=device, **kw)
    try:
        return a, b, ts.run(W.series_params(b - a, total=total, first=a), p)
    finally:
        ts.close()
