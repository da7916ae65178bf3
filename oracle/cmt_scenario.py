"""A small CMT-inversion scenario for the UNMODIFIED libs/loc_cmt_util.py, set up exactly as the reference notebook
does (source_location_cmt_W2L2_Figs_9_10_11.ipynb cells 23-49), with oracle/pyprop8_stub.py standing in for the absent
pyprop8.  Test infrastructure: tests/golden/make_golden.py runs it over the unmodified reference classes (CPU, build
container) and tests/test_gpu_dropin.py over the B200 shim; both call `cmt_util.optfunc_OT` (libs/loc_cmt_util.py:186-306).
"""
import numpy as np

RECX = np.array([60.0, -50.0, 30.0, -80.0])
RECY = np.array([40.0, 70.0, -90.0, -30.0])
TRUE = (10.0, 5.0, 15.0)
MO = 1.0e13


def build_optdata(cmt_util, cmt=False, wopt="Wavg"):
    """-> (optdata, t): the dictionaries of notebook cells 23, 29, 34, 46, 48 (no preconditioning)."""
    from libs import loc_cmt_util_opt
    loc_cmt_util_opt.init()
    prop8data = {"model": None, "sdrm": [302, 88, -14, MO], "recx": RECX, "recy": RECY}            # cell 23
    t, clean = cmt_util.prop8seis(TRUE[0], TRUE[1], TRUE[2], prop8data, show_progress=False)
    nr, nc, nt = clean.shape
    ph = 0.7 * np.arange(nr * nc).reshape(nr, nc, 1)
    noise = 0.05 * np.abs(clean).max() * np.sin(0.37 * t[None, None, :] + ph)                      # cell 29 (deterministic)
    prop8data["obs_seis"] = clean + noise
    invopt, OTdata = {}, {}                                                                         # cell 34
    invopt["loc"], invopt["cmt"], invopt["mistype"], invopt["precon"] = True, bool(cmt), "OT", False
    OTdata["plambda"] = OTdata["olambda"] = 0.04
    OTdata["distfunc"], OTdata["Wopt"], OTdata["theta"] = "W2", wopt, 45.0
    obs_grids = cmt_util.buildFingerprintwindows(t, prop8data["obs_seis"])                          # cell 46
    OTdata["obs_grids01"] = cmt_util.buildFingerprintwindows(t, prop8data["obs_seis"], u0=0.0, u1=1.0)
    wfobs, wfobs_target = cmt_util.BuildOTobjfromWaveform(t, prop8data["obs_seis"], obs_grids, OTdata,
                                                          lambdav=OTdata["olambda"], theta=OTdata["theta"])
    OTdata["wfobs"], OTdata["wfobs_target"], OTdata["obs_grids"] = wfobs, wfobs_target, obs_grids
    invopt["mref"] = list(TRUE)
    invopt["mscal"] = np.ones(9 if cmt else 3)                                                      # cell 48, no precon
    return {"invopt": invopt, "OTdata": OTdata, "prop8data": prop8data}, t


def trial_models(cmt_util, cmt=False):
    locs = [np.array([40.0, 40.0, 10.0]), np.array([12.0, 3.0, 18.0]), np.array([-20.0, 30.0, 25.0])]   # cell 42: first one
    if not cmt:
        return locs
    from pyprop8.utils import make_moment_tensor, rtf2xyz
    M = rtf2xyz(make_moment_tensor(302, 88, -14, MO * 1.0e-13, 0, 0))
    up = M[np.triu_indices(3)]
    return [np.append(l, up * (1.0 + 0.15 * np.cos(np.arange(6) + i))) for i, l in enumerate(locs)]
