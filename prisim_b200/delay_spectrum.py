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


def window_N2width(n_window=None, shape="rect", fftpow=1.0, area_normalize=True, power_normalize=False):
    """Width of the equivalent rectangular window as a fraction of the window length (``DSP.window_N2width`` of the
    un-vendored astroutils [AU-memory], used at interferometry.py:8236 and delay_spectrum.py:2155): sum(w / max w) / N
    (rect 1.0, bhw 0.35875, bnw 0.3635819), or with power_normalize sum((w / max w)^2) / N (bhw 0.2580); evaluated on
    a long window when n_window is None."""
    n = 1000000 if n_window is None else int(n_window)
    w = windowing(n, shape=shape.lower()) ** fftpow
    w = w / w.max()
    return float(NP.sum(w ** 2) / n) if (power_normalize and not area_normalize) else float(NP.sum(w) / n)


def window_fftpow(N_window, shape="rect", fftpow=1.0, centering=True, peak=None, area_normalize=False, power_normalize=False):
    """``DSP.window_fftpow`` of astroutils for fftpow = 1 (the default of every PRISim call site): the plain window.
    The construction for fftpow != 1 (window whose Fourier transform is raised to a power) lives in the un-vendored
    astroutils and could not be pinned here, so it is refused rather than guessed."""
    if float(fftpow) != 1.0:
        raise NotImplementedError("window_fftpow with fftpow != 1 is defined in the un-vendored astroutils package and is not restated here")
    return windowing(int(N_window), shape=shape.lower(), centering=centering, area_normalize=area_normalize, power_normalize=power_normalize,
                     peak=1.0 if peak is None else peak)


def subband_weights(f, df, bw_eff, freq_center, shape="rect", fftpow=1.0):
    """The [n_win, nchan] frequency weights of ``DelaySpectrum.subband_delay_transform`` (delay_spectrum.py:2153-2176):
    for every (effective bandwidth, centre frequency) a power-normalised window of n = round(bw_eff / (frac_width df))
    samples scaled by sqrt(frac_width n), centred on the channel nearest to the centre frequency (within df/2; others are
    dropped), clipped to the band; windows ordered by centre channel.  Returns (weights, kept centre-channel indices)."""
    f = NP.asarray(f, dtype=NP.float64)
    bw_eff = NP.asarray(bw_eff, dtype=NP.float64).reshape(-1)
    freq_center = NP.asarray(freq_center, dtype=NP.float64).reshape(-1)
    frac_width = window_N2width(n_window=None, shape=shape, fftpow=fftpow, area_normalize=False, power_normalize=True)
    n_window = NP.round(bw_eff / frac_width / df).astype(int)                           # :2156-2157
    nearest = NP.rint((freq_center - f[0]) / df).astype(int)                             # LKP.find_1NN within df/2, out-of-band removed
    ok = (nearest >= 0) & (nearest < f.size)
    ok &= NP.abs(f[NP.clip(nearest, 0, f.size - 1)] - freq_center) <= 0.5 * df
    nearest, n_window = nearest[ok], n_window[ok]
    order = NP.argsort(nearest, kind="stable")                                            # :2159-2163
    nearest, n_window = nearest[order], n_window[order]
    wts = NP.zeros((nearest.size, f.size))
    for i, (c, n) in enumerate(zip(nearest, n_window)):
        win = NP.sqrt(frac_width * n) * window_fftpow(n, shape=shape, fftpow=fftpow, centering=True, power_normalize=True)   # :2166
        k = c + NP.arange(n) - int(n / 2)                                                 # :2168
        inside = (k >= 0) & (k < f.size)
        wts[i, k[inside]] = win[inside]
    return wts, nearest


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
        self.cc_lags = None                                              # no delay CLEAN here: the 'cc' branch of the sub-band transform is skipped (:2151)
        self.subband_delay_spectra = {}
        self.subband_delay_spectra_resampled = {}

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

    def subband_delay_transform(self, bw_eff, freq_center=None, shape=None, fftpow=None, pad=None, bpcorrect=False, action=None,
                                verbose=True):
        """Delay transforms on frequency sub-bands, same call as delay_spectrum.py:1842-2248.  Arguments are dictionaries
        with the keys 'cc' and 'sim' like the reference's; only the 'sim' branch is computed (the 'cc' branch needs delay
        CLEAN products, which are outside the hot-path scope -- the reference skips it too while ``cc_lags`` is None,
        :2151).  Stores ``subband_delay_spectra`` (full resolution: 'skyvis_lag', 'vis_lag', 'vis_noise_lag', 'lag_kernel' of
        shape [nbl, n_win, nchan + npad, n_t], 'freq_wts', 'lags', 'lag_corr_length', ...) and
        ``subband_delay_spectra_resampled`` (:2220-2240: decimated by min((nchan + npad) df / bw_eff); kernel and lags by
        linear interpolation, spectra by Fourier resampling) and returns one of them for action = 'return_oversampled' /
        'return_resampled'.  Every (window, snapshot) is one launch of the delay-transform kernel with the window as its
        weights; the resampling of the returned numpy arrays is host post-processing (scipy.signal.resample)."""
        if not isinstance(bw_eff, dict):
            raise TypeError("Effective bandwidth must be specified as a dictionary")
        bw = {}
        for key in ("cc", "sim"):
            if key in bw_eff:
                if not isinstance(bw_eff[key], (int, float, list, NP.ndarray)):
                    raise TypeError("Value of effective bandwidth must be a scalar, list or numpy array")
                bw[key] = NP.asarray(bw_eff[key], dtype=NP.float64).reshape(-1)
                if NP.any(bw[key] <= 0.0):
                    raise ValueError("All values in effective bandwidth must be strictly positive")
        if "sim" not in bw:
            raise KeyError("Effective bandwidth for key 'sim' must be specified")
        if freq_center is None:
            fc = {key: NP.asarray(self.f[self.f.size // 2]).reshape(-1) for key in bw}
        elif isinstance(freq_center, dict):
            fc = {}
            for key in bw:
                if not isinstance(freq_center.get(key, None), (int, float, list, NP.ndarray)):
                    raise TypeError("Values(s) of frequency center must be scalar, list or numpy array")
                fc[key] = NP.asarray(freq_center[key], dtype=NP.float64).reshape(-1)
                if NP.any((fc[key] <= self.f.min()) | (fc[key] >= self.f.max())):
                    raise ValueError("Value(s) of frequency center(s) must lie strictly inside the observing band")
        else:
            raise TypeError("Input frequency center must be specified as a dictionary")
        for key in bw:
            if bw[key].size == 1 and fc[key].size > 1:
                bw[key] = NP.repeat(bw[key], fc[key].size)
            elif bw[key].size > 1 and fc[key].size == 1:
                fc[key] = NP.repeat(fc[key], bw[key].size)
            elif bw[key].size != fc[key].size:
                raise ValueError("Effective bandwidth(s) and frequency center(s) must have same number of elements")
        if shape is not None:
            if not isinstance(shape, dict):
                raise TypeError("Window shape must be specified as a dictionary")
            for key in bw:
                if not isinstance(shape[key], str):
                    raise TypeError("Window shape must be a string")
                if shape[key] not in ["rect", "bhw", "bnw", "RECT", "BHW", "BNW"]:
                    raise ValueError("Invalid value for window shape specified.")
        else:
            shape = {key: "rect" for key in bw}
        if fftpow is None:
            fftpow = {key: 1.0 for key in bw}
        else:
            if not isinstance(fftpow, dict):
                raise TypeError("Power to raise FFT of window by must be specified as a dictionary")
            for key in bw:
                if not isinstance(fftpow[key], (int, float)):
                    raise TypeError("Power to raise window FFT by must be a scalar value.")
                if fftpow[key] < 0.0:
                    raise ValueError("Power for raising FFT of window by must be positive.")
        if pad is None:
            pad = {key: 1.0 for key in bw}
        else:
            if not isinstance(pad, dict):
                raise TypeError("Padding for delay transform must be specified as a dictionary")
            pad = dict(pad)
            for key in bw:
                if not isinstance(pad[key], (int, float)):
                    raise TypeError("pad fraction must be a scalar value.")
                if pad[key] < 0.0:
                    pad[key] = 0.0
        if not isinstance(bpcorrect, bool):
            raise TypeError("Input keyword bpcorrect must be of boolean type")

        import torch
        ia = self.ia
        key = "sim"
        nbl, nchan, nsnap = ia.baselines.shape[0], self.f.size, len(ia._skyvis)
        freq_wts, _ = subband_weights(self.f, self.df, bw[key], fc[key], shape=shape[key], fftpow=fftpow[key])
        nwin = freq_wts.shape[0]
        npad = int(nchan * pad[key])                                                          # :2178
        lags = NP.fft.fftshift(NP.fft.fftfreq(nchan + npad, d=self.df))                        # :2179
        wts_dev = engine._f64(freq_wts, ia.device)
        res = {"freq_center": fc[key], "shape": shape[key], "freq_wts": freq_wts, "bw_eff": bw[key], "npad": npad, "lags": lags,
               "lag_corr_length": nchan / NP.sum(freq_wts, axis=1)}
        for name, lst in (("skyvis_lag", ia._skyvis), ("vis_lag", ia._vis), ("vis_noise_lag", ia._noise), ("lag_kernel", None)):
            if lst is not None and not lst:
                continue                                                                      # product not generated (the reference needs all three)
            out = torch.empty((nbl, nwin, nchan + npad, nsnap), dtype=torch.complex128, device=ia._dev_str())
            for t in range(nsnap):
                for i in range(nwin):
                    if lst is None:
                        k = engine.delay_transform(None, ia._bp[t], wts_dev[i], self.df, pad=pad[key], downsample=False,
                                                   nrows=nbl if ia._bp[t].ndim == 2 else 1, nchan=nchan, device=ia.device)
                    else:
                        k = engine.delay_transform(lst[t], ia._bp[t], wts_dev[i], self.df, pad=pad[key], downsample=False)
                    out[:, i, :, t] = k
            res[name] = out.cpu().numpy()
        result = {key: res}
        self.subband_delay_spectra = result
        # resampled products (:2220-2240)
        from scipy import signal
        rs = {"freq_center": res["freq_center"], "bw_eff": res["bw_eff"]}
        factor = float(NP.min((nchan + npad) * self.df / rs["bw_eff"]))
        pos = NP.arange(0, lags.size, factor)
        rs["lags"] = NP.interp(pos, NP.arange(lags.size), lags)

        def interp_axis2(a):                    # DSP.downsampler(method='interp', kind='linear') along axis 2
            j0 = NP.floor(pos).astype(int)
            fr = (pos - j0).reshape(1, 1, -1, 1)
            j1 = NP.minimum(j0 + 1, a.shape[2] - 1)
            out = a[:, :, j0, :] * (1.0 - fr) + a[:, :, j1, :] * fr
            out[:, :, pos > a.shape[2] - 1, :] = NP.nan          # beyond the last sample: interp1d(bounds_error=False) fills NaN
            return out

        rs["lag_kernel"] = interp_axis2(res["lag_kernel"])
        nout = int(NP.round(lags.size / factor))
        for name in ("skyvis_lag", "vis_lag", "vis_noise_lag"):
            if name in res:
                rs[name] = signal.resample(res[name], nout, axis=2)                           # DSP.downsampler(method='FFT') [AU-memory]
        dlag = rs["lags"][1] - rs["lags"][0]
        rs["lag_corr_length"] = (1.0 / res["bw_eff"]) / dlag
        self.subband_delay_spectra_resampled = {key: rs}
        if action == "return_oversampled":
            return result
        if action == "return_resampled":
            return self.subband_delay_spectra_resampled
