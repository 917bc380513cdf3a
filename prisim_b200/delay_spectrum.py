"""Mirror of ``prisim/delay_spectrum.py`` for ``DelaySpectrum.delay_transform``.

Wraps an ``InterferometerArray`` like the reference (delay_spectrum.py:901, :1178-1188) and runs
the same GPU kernel as ``InterferometerArray.delay_transform``; the dictionary it returns has the
reference's keys (delay_spectrum.py:1302-1342).  Delay CLEAN, subband transforms and the power
spectrum classes are outside the hot-path scope.
"""
from __future__ import annotations

import numpy as NP

from . import engine
from .interferometry import InterferometerArray


def windowing(N, shape="rect", pad_width=0, centering=True, area_normalize=False, peak=1.0, power_normalize=False):
    """Host helper producing the spectral window run_prisim passes as ``freq_wts``
    (scripts/run_prisim.py:954: nchan * windowing(nchan, 'bhw', area_normalize=True)).  Semantics
    of the un-vendored ``astroutils.DSP_modules.windowing``: 'rect', 4-term Blackman-Harris 'bhw'
    or Blackman-Nuttall 'bnw' over n/(N-1)."""
    n = NP.arange(N)
    if shape == "rect":
        win = NP.ones(N)
    elif shape in ("bhw", "bnw"):
        a = {"bhw": (0.35875, 0.48829, 0.14128, 0.01168), "bnw": (0.3635819, 0.4891775, 0.1365995, 0.0106411)}[shape]
        x = 2 * NP.pi * n / (N - 1)
        win = a[0] - a[1] * NP.cos(x) + a[2] * NP.cos(2 * x) - a[3] * NP.cos(3 * x)
    else:
        raise ValueError("Window shape must be 'rect', 'bhw' or 'bnw'")
    if area_normalize:
        win = win / NP.sum(win)
    elif power_normalize:
        win = win / NP.sqrt(NP.sum(win ** 2))
    else:
        win = win * peak / NP.amax(win)
    if pad_width > 0:
        win = NP.pad(win, (pad_width, pad_width), mode="constant")
    return win


def window_N2width(n_window=None, shape="rect", fftpow=1.0):
    """Width of the equivalent rectangular window as a fraction of the window length (``DSP.window_N2width`` of the
    un-vendored astroutils, used at interferometry.py:8236): sum(w / max w) / N, evaluated on a long window when
    n_window is None (rect 1.0, bhw 0.35875, bnw 0.3635819)."""
    n = 1000000 if n_window is None else int(n_window)
    w = windowing(n, shape=shape.lower()) ** fftpow
    return float(NP.sum(w / w.max()) / n)


class DelaySpectrum(object):
    def __init__(self, interferometer_array=None, init_file=None):
        if init_file is not None:
            raise NotImplementedError("saved delay spectra are outside the hot-path scope")
        if not isinstance(interferometer_array, InterferometerArray):
            raise TypeError("Input interferometer_array must be an instance of class InterferometerArray")
        self.ia = interferometer_array                                   # delay_spectrum.py:1178
        self.f = self.ia.channels
        self.df = self.ia.freq_resolution
        self.n_acc = self.ia.n_acc
        self.horizon_delay_limits = self.get_horizon_delay_limits() if self.ia.n_acc > 0 else None   # :1183
        self.pad = None
        self.lags = None
        self.bp_wts = None
        self.vis_lag = self.skyvis_lag = self.vis_noise_lag = self.lag_kernel = None

    @property
    def bp(self):
        return self.ia.bp

    def get_horizon_delay_limits(self, phase_center=None, phase_center_coords=None):
        """delay_spectrum.py:2976-3030: [n_phase_centres, nbl, 2] minimum / maximum horizon delay of every baseline."""
        from . import baseline_delay_horizon as DLY
        from . import geometry as GEOM
        if phase_center is None:
            phase_center = self.ia.phase_center
            phase_center_coords = self.ia.phase_center_coords
        if phase_center_coords not in ["hadec", "altaz", "dircos"]:
            raise ValueError('Phase center coordinates must be "altaz", "hadec" or "dircos"')
        if phase_center_coords == "hadec":
            pc_dircos = GEOM.altaz2dircos(GEOM.hadec2altaz(phase_center, self.ia.latitude, units="degrees"), units="degrees")
        elif phase_center_coords == "altaz":
            pc_dircos = GEOM.altaz2dircos(phase_center, units="degrees")
        else:
            pc_dircos = phase_center
        return DLY.horizon_delay_limits(self.ia.baselines, pc_dircos, units="mks")

    def delay_transform_allruns(self, vis, pad=1.0, freq_wts=None, downsample=True, verbose=True):
        """delay_spectrum.py:1475-1618: delay-transform external visibilities of shape (..., nbl, nchan, ntimes) (several
        runs / realisations) with this object's bandpass; returns {'freq_wts', 'pad', 'lags', 'vis_lag', 'lag_kernel'}
        with the reference's shapes.  Every (run, snapshot) slice is one launch of the delay-transform kernel."""
        if not isinstance(vis, NP.ndarray):
            raise TypeError("Input vis must be a numpy array")
        ia = self.ia
        nbl, nchan, nt = ia.baselines.shape[0], self.f.size, self.n_acc
        if vis.ndim < 3:
            raise ValueError("Input vis must be at least 3-dimensional")
        if vis.shape[-3:] != (nbl, nchan, nt):
            raise ValueError("Input vis does not have compatible shape")
        if vis.ndim == 3:
            vis = vis.reshape((1,) + vis.shape)
        if not isinstance(pad, (int, float)):
            raise TypeError("pad fraction must be a scalar value.")
        if pad < 0.0:
            pad = 0.0
        if not isinstance(downsample, bool):
            raise TypeError("Input downsample must be of boolean type")
        lead = vis.shape[:-3]
        ones = (1,) * len(lead)
        if freq_wts is not None:                                                          # :1541-1551
            freq_wts = NP.asarray(freq_wts)
            if freq_wts.shape == self.f.shape:
                wts_full = freq_wts.reshape(ones + (1, -1, 1))
            elif freq_wts.shape == (nchan, nt):
                wts_full = freq_wts.reshape(ones + (1, nchan, nt))
            elif freq_wts.shape == (nbl, nchan):
                wts_full = freq_wts.reshape(ones + (nbl, nchan, 1))
            elif freq_wts.shape == (nbl, nchan, nt):
                wts_full = freq_wts.reshape(ones + (nbl, nchan, nt))
            elif freq_wts.shape == vis.shape:
                raise NotImplementedError("per-run frequency weights are not on the device path")
            else:
                raise ValueError("window shape dimensions incompatible with number of channels and/or number of tiemstamps.")
            getter = ia._freq_wts_getter(freq_wts)
        else:                                                                             # :1552-1553 the stored weights
            stored = self.bp_wts if self.bp_wts is not None else ia.bp_wts
            wts_full = NP.asarray(stored).reshape(ones + NP.asarray(stored).shape)
            getter = ia._bp_wts if self.bp_wts is None else ia._freq_wts_getter(self.bp_wts)
        nout = engine.delay_nout(nchan, pad, downsample)
        flat = vis.reshape((-1, nbl, nchan, nt))
        out = NP.empty((flat.shape[0], nbl, nout, nt), dtype=NP.complex128)
        kern = NP.empty((nbl, nout, nt), dtype=NP.complex128)
        for t in range(nt):
            bp = ia._bp[t]
            wts = None if getter is None else getter(t)
            krows = nbl if (bp.ndim == 2 or (wts is not None and wts.ndim == 2)) else 1
            k = engine.delay_transform(None, bp, wts, self.df, pad=pad, downsample=downsample, nrows=krows, nchan=nchan, device=ia.device)
            kern[:, :, t] = k.cpu().numpy()
            for r in range(flat.shape[0]):
                x = engine._c128(NP.ascontiguousarray(flat[r, :, :, t]), ia.device)
                out[r, :, :, t] = engine.delay_transform(x, bp, wts, self.df, pad=pad, downsample=downsample).cpu().numpy()
        lags = NP.fft.fftshift(NP.fft.fftfreq(int(nchan * (1 + pad)), d=self.df))           # :1575
        if downsample and pad > 0.0:                                                          # :1596
            lags = NP.interp(NP.arange(0, lags.size, 1 + pad), NP.arange(lags.size), lags)
        return {"freq_wts": wts_full, "pad": pad, "lags": lags.flatten(), "vis_lag": out.reshape(lead + (nbl, nout, nt)),
                "lag_kernel": kern.reshape(ones + (nbl, nout, nt))}

    def delay_transform(self, pad=1.0, freq_wts=None, downsample=True, action=None, verbose=True):
        """delay_spectrum.py:1224-1342."""
        if not isinstance(pad, (int, float)):
            raise TypeError("pad fraction must be a scalar value.")
        if pad < 0.0:
            pad = 0.0
        if not isinstance(downsample, bool):
            raise TypeError("Input downsample must be of boolean type")
        ia = self.ia
        nbl, nchan = ia.baselines.shape[0], ia.channels.size
        getter = ia._bp_wts if freq_wts is None else ia._freq_wts_getter(freq_wts)
        lags = NP.fft.fftshift(NP.fft.fftfreq(int(nchan * (1 + pad)), d=self.df))           # :1305
        out = {"vis_lag": [], "skyvis_lag": [], "vis_noise_lag": [], "lag_kernel": []}
        src = {"vis_lag": ia._vis, "skyvis_lag": ia._skyvis, "vis_noise_lag": ia._noise}
        for t in range(len(ia._skyvis)):
            bp = ia._bp[t]
            wts = None if getter is None else getter(t)
            for key, lst in src.items():
                if lst:
                    out[key].append(engine.delay_transform(lst[t], bp, wts, self.df, pad=pad, downsample=downsample))
            krows = nbl if (bp.ndim == 2 or (wts is not None and wts.ndim == 2)) else 1
            kern = engine.delay_transform(None, bp, wts, self.df, pad=pad, downsample=downsample, nrows=krows,
                                          nchan=nchan, device=ia.device)
            out["lag_kernel"].append(kern if krows == nbl else kern[0])
        if downsample and pad > 0.0:                                                          # :1327
            pos = NP.arange(0, lags.size, 1 + pad)
            lags = NP.interp(pos, NP.arange(lags.size), lags)
        result = {"pad": pad, "lags": lags,
                  "freq_wts": (NP.ones((nbl, nchan, len(ia._skyvis))) if getter is None else
                               ia._stack([getter(t) for t in range(len(ia._skyvis))], expand=True))}
        for key, lst in out.items():
            result[key] = ia._stack(lst, expand=(key == "lag_kernel")) if lst else None
        if action == "store":                                                                  # :1333-1340
            self.pad = pad
            self.lags = result["lags"]
            self.bp_wts = result["freq_wts"]
            self.vis_lag, self.skyvis_lag = result["vis_lag"], result["skyvis_lag"]
            self.vis_noise_lag, self.lag_kernel = result["vis_noise_lag"], result["lag_kernel"]
        return result
