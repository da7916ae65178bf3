"""Stand-in for the third-party package `pyprop8`, which the reference's CMT adapters import at module scope
(libs/loc_cmt_util.py:9,12) and which is neither vendored in /root/reference nor installed in this image.

Test infrastructure only (like the rest of oracle/): it lets the UNMODIFIED libs/loc_cmt_util.py be imported and its
`optfunc_OT` (libs/loc_cmt_util.py:186-306) be driven end to end - once over the unmodified reference classes in the
build container (tests/golden/make_golden.py -> tests/golden/cmt_optfunc.npz) and once over the B200 shim on the GPU
box (tests/test_gpu_dropin.py).  It is NOT a seismogram code: `compute_seismograms` returns a smooth, deterministic,
pyprop8-SHAPED synthetic (far-field pulse of a moment-tensor point source in a uniform medium) with the array layouts
loc_cmt_util expects:

    t (nt,), seismograms (nstations, 3, nt), derivatives (nstations, nderivs, 3, nt)

with the derivative slots laid out as `DerivativeSwitches` announces them (i_mt .. i_mt+5 in pyprop8's diagonal-first
order Mxx, Myy, Mzz, Mxy, Mxz, Myz; i_x, i_y, i_z).  Only what loc_cmt_util touches is provided.
"""
import sys
import types

import numpy as np

_V = 3.2          # km/s: puts the arrivals of 40-160 km paths inside a 61 s window
_A = 0.22         # 1/s: pulse width (a ~10 s Ricker-like pulse, cf. the reference's 0.05-0.2 Hz band)


class PointSource:
    def __init__(self, x, y, dep, Mxyz, F, time):
        self.x, self.y, self.dep = float(x), float(y), float(dep)
        self.Mxyz = np.asarray(Mxyz, dtype=np.float64).reshape(1, 3, 3)
        self.F = F
        self.time = time


class ListOfReceivers:
    def __init__(self, xx, yy, depth=0.0, geometry="cartesian"):
        self.xx = np.asarray(xx, dtype=np.float64).reshape(-1)
        self.yy = np.asarray(yy, dtype=np.float64).reshape(-1)
        self.depth = depth
        self.nstations = self.xx.size
        self.rr = np.zeros(self.nstations)
        self.pp = np.zeros(self.nstations)


class DerivativeSwitches:
    def __init__(self, moment_tensor=False, force=False, r=False, phi=False, x=False, y=False, z=False,
                 time=False, thickness=False, structure=None):
        if r or phi or force or time or thickness:
            raise NotImplementedError("pyprop8 stand-in: only x, y, z and moment_tensor derivatives")
        self.moment_tensor, self.x, self.y, self.z = bool(moment_tensor), bool(x), bool(y), bool(z)
        self.r = self.phi = False
        n = 0
        self.i_mt = self.i_x = self.i_y = self.i_z = self.i_r = self.i_phi = None
        if moment_tensor:
            self.i_mt = n; n += 6
        if x:
            self.i_x = n; n += 1
        if y:
            self.i_y = n; n += 1
        if z:
            self.i_z = n; n += 1
        self.nderivs = n


def _pulse(tau):
    a2 = (_A * tau) ** 2
    return (1.0 - 2.0 * a2) * np.exp(-a2)


def _seis(sx, sy, sz, M, stations, t):
    dx, dy = stations.xx - sx, stations.yy - sy
    r = np.sqrt(dx * dx + dy * dy + sz * sz)
    gam = np.stack([dx / r, dy / r, -sz / r * np.ones_like(r)], axis=1)          # (nr, 3) direction cosines
    rad = np.einsum("ia,ab,ib->i", gam, M, gam)                                   # gamma^T M gamma
    g = _pulse(t[None, :] - (r / _V)[:, None] - 6.0)                             # (nr, nt)
    return (gam * (rad / r)[:, None] * 40.0)[:, :, None] * g[:, None, :], gam, r, g


def compute_seismograms(model, source, stations, nt, timestep, alpha, source_time_function=None,
                        derivatives=None, show_progress=True, **kw):
    t = np.arange(nt) * float(timestep)
    M = source.Mxyz[0]
    s, gam, r, g = _seis(source.x, source.y, source.dep, M, stations, t)
    stations.rr = np.sqrt((stations.xx - source.x) ** 2 + (stations.yy - source.y) ** 2)
    stations.pp = np.arctan2(stations.yy - source.y, stations.xx - source.x)
    if derivatives is None:
        return t, s
    drv = derivatives
    d = np.zeros((stations.nstations, drv.nderivs, 3, nt))
    if drv.moment_tensor:                       # linear in M: diagonal-first order, off-diagonals count twice (symmetric M)
        pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
        for k, (a, b) in enumerate(pairs):
            w = gam[:, a] * gam[:, b] * (1.0 if a == b else 2.0)
            d[:, drv.i_mt + k] = (gam * (w / r)[:, None] * 40.0)[:, :, None] * g[:, None, :]
    h = 1e-4                                    # km: central differences of the smooth synthetic
    for on, slot, e in ((drv.x, drv.i_x, (h, 0, 0)), (drv.y, drv.i_y, (0, h, 0)), (drv.z, drv.i_z, (0, 0, h))):
        if on:
            sp = _seis(source.x + e[0], source.y + e[1], source.dep + e[2], M, stations, t)[0]
            sm = _seis(source.x - e[0], source.y - e[1], source.dep - e[2], M, stations, t)[0]
            d[:, slot] = (sp - sm) / (2.0 * h)
    return t, s, d


# ---- pyprop8.utils (libs/loc_cmt_util.py:12)
def make_moment_tensor(strike, dip, rake, M0, eta, xtr):
    """Double couple in (r, theta, phi) (Aki & Richards 4.97 with x = north); eta / xtr (CLVD / isotropic parts) unused."""
    s, d, l = np.deg2rad(strike), np.deg2rad(dip), np.deg2rad(rake)
    Mxx = -M0 * (np.sin(d) * np.cos(l) * np.sin(2 * s) + np.sin(2 * d) * np.sin(l) * np.sin(s) ** 2)
    Mxy = M0 * (np.sin(d) * np.cos(l) * np.cos(2 * s) + 0.5 * np.sin(2 * d) * np.sin(l) * np.sin(2 * s))
    Mxz = -M0 * (np.cos(d) * np.cos(l) * np.cos(s) + np.cos(2 * d) * np.sin(l) * np.sin(s))
    Myy = M0 * (np.sin(d) * np.cos(l) * np.sin(2 * s) - np.sin(2 * d) * np.sin(l) * np.cos(s) ** 2)
    Myz = -M0 * (np.cos(d) * np.cos(l) * np.sin(s) - np.cos(2 * d) * np.sin(l) * np.cos(s))
    Mzz = M0 * np.sin(2 * d) * np.sin(l)
    # north-east-down -> (r, theta, phi) = (up, south, east)
    return np.array([[Mzz, Mxz, -Myz], [Mxz, Mxx, -Mxy], [-Myz, -Mxy, Myy]])


def rtf2xyz(M):
    """(r, theta, phi) = (up, south, east) -> (x, y, z) = (east, north, up)."""
    M = np.asarray(M, dtype=np.float64)
    P = np.array([[0.0, 0.0, 1.0], [0.0, -1.0, 0.0], [1.0, 0.0, 0.0]])
    return P @ M @ P.T


def stf_trapezoidal(omega, trise, trupt):
    return np.ones_like(np.asarray(omega, dtype=np.float64))


def clp_filter(w, w0, w1):
    return np.ones_like(np.asarray(w, dtype=np.float64))


def install():
    """Register this module as `pyprop8` / `pyprop8.utils` unless the real package is importable."""
    try:
        import pyprop8  # noqa: F401
        return False
    except Exception:
        pass
    me = sys.modules[__name__]
    pkg = types.ModuleType("pyprop8")
    for k in ("PointSource", "ListOfReceivers", "DerivativeSwitches", "compute_seismograms"):
        setattr(pkg, k, getattr(me, k))
    pkg.__path__ = []
    utils = types.ModuleType("pyprop8.utils")
    for k in ("rtf2xyz", "make_moment_tensor", "stf_trapezoidal", "clp_filter"):
        setattr(utils, k, getattr(me, k))
    pkg.utils = utils
    sys.modules["pyprop8"] = pkg
    sys.modules["pyprop8.utils"] = utils
    return True
