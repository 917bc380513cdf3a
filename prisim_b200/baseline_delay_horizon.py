"""Mirror of ``prisim/baseline_delay_horizon.py`` for the functions on the hot path.

``geometric_delay`` (baseline_delay_horizon.py:133-241) returns the dense [nsrc, nbl] delay matrix.
The visibility kernel never materialises it (it forms s.b/c per (source, baseline) in registers);
this standalone version exists for API parity and diagnostics: the coordinate conversion runs in
the ``pb200_sky_cull`` kernel and the dense product is one fp64 library GEMM.
"""
from __future__ import annotations

import numpy as NP
import scipy.constants as FCNST
import torch

from . import engine


def geometric_delay(baselines, skypos, altaz=False, dircos=False, hadec=True, units="mks", latitude=None, device=None):
    try:
        baselines, skypos
    except NameError:
        raise NameError("baselines and/or skypos not defined in geometric_delay().")
    if (altaz) + (dircos) + (hadec) != 1:
        raise ValueError("One and only one of altaz, dircos, hadec must be set to True.")   # :178-179
    if hadec and (latitude is None):
        raise ValueError("Latitude must be specified when skypos is in HA-Dec format.")      # :181-182
    baselines = NP.asarray(baselines, dtype=NP.float64)
    if baselines.ndim == 1:
        baselines = baselines.reshape(1, -1)
    if baselines.shape[1] < 3:
        baselines = NP.hstack((baselines, NP.zeros((baselines.shape[0], 3 - baselines.shape[1]))))
    elif baselines.shape[1] > 3:
        baselines = baselines[:, :3]
    skypos = NP.asarray(skypos, dtype=NP.float64)
    ncol = 3 if dircos else 2
    if skypos.ndim < 2:
        if skypos.size != ncol:
            raise ValueError("Sky position should consist of {0} elements.".format(ncol))
        skypos = skypos.reshape(1, -1)
    elif skypos.ndim > 2 or skypos.shape[1] != ncol:
        raise ValueError("Sky positions should be a Nx{0} numpy array.".format(ncol))
    device = engine._dev(device)
    coords = "altaz" if altaz else ("dircos" if dircos else "hadec")
    dc, _ = engine.sky_cull(skypos, coords, latitude_deg=0.0 if latitude is None else latitude, roi_radius_deg=180.0,
                            device=device)
    c = FCNST.c if units == "mks" else FCNST.c * 1e2
    bl = engine._f64(baselines, device)
    return (torch.matmul(dc, bl.T) / c).cpu().numpy()                                        # :240


def horizon_delay_limits(baselines, refdir, units="mks"):
    """baseline_delay_horizon.py:100-129: per-baseline [min, max] horizon delays about the phase
    centre `refdir` (direction cosines).  Returns [nref, nbl, 2]."""
    baselines = NP.asarray(baselines, dtype=NP.float64).reshape(-1, 3)
    refdir = NP.asarray(refdir, dtype=NP.float64).reshape(-1, 3)
    c = FCNST.c if units == "mks" else FCNST.c * 1e2
    blen = NP.sqrt(NP.sum(baselines ** 2, axis=1)) / c
    off = NP.dot(refdir, baselines.T) / c
    return NP.stack((-blen[NP.newaxis, :] - off, blen[NP.newaxis, :] - off), axis=2)
