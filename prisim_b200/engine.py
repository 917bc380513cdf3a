"""Thin typed wrappers over the C-ABI: torch CUDA tensors in, torch CUDA tensors out.

Each function corresponds to one entry point of ``include/prisim_b200.h`` and, through it, to the
reference lines cited there.  No numerical work happens in Python here.
"""
from __future__ import annotations

import ctypes as C

import numpy as NP
import torch

from . import _lib
from ._lib import BeamDesc, SpectrumDesc, _ptr, get_context


def _dev(device):
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else 0
    return int(device)


def _f64(x, device):
    """Host array-like or tensor -> contiguous float64 CUDA tensor on `device`."""
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda:{0}".format(device), dtype=torch.float64).contiguous()
    return torch.as_tensor(NP.ascontiguousarray(x, dtype=NP.float64)).to("cuda:{0}".format(device))


def _c128(x, device):
    """Host array-like or tensor -> contiguous complex128 CUDA tensor on `device`."""
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda:{0}".format(device), dtype=torch.complex128).contiguous()
    return torch.as_tensor(NP.ascontiguousarray(x, dtype=NP.complex128)).to("cuda:{0}".format(device))


def _h64(x):
    return NP.ascontiguousarray(NP.asarray(x, dtype=NP.float64))


def sky_cull(skypos, coords, latitude_deg=0.0, roi_radius_deg=90.0, roi_center_dircos=None, device=None):
    """``pb200_sky_cull``: returns (dircos [nsrc,3] f64, index [nsrc] i32) of the sources kept.
    Replaces interferometry.py:6174-6219."""
    device = _dev(device)
    ctx = get_context(device)
    skypos = _f64(skypos, device)
    nsrc0 = skypos.shape[0] if skypos.ndim == 2 else 0
    code = {"altaz": _lib.SKY_ALTAZ, "hadec": _lib.SKY_HADEC, "dircos": _lib.SKY_DIRCOS}[coords]
    dircos = torch.empty((max(nsrc0, 1), 3), dtype=torch.float64, device=skypos.device)
    index = torch.empty((max(nsrc0, 1),), dtype=torch.int32, device=skypos.device)
    n = C.c_int(0)
    center = None if roi_center_dircos is None else _h64(roi_center_dircos)
    ctx.check(ctx.lib.pb200_sky_cull(ctx.handle, _ptr(skypos), nsrc0, code, float(latitude_deg),
                                     float(roi_radius_deg), _ptr(center), _ptr(dircos), _ptr(index),
                                     C.byref(n), ctx.stream()))
    return dircos[: n.value], index[: n.value]


def make_beam_desc(**kw):
    """Fill a ``pb200_beam_desc``; device arrays for the element array are kept alive on the
    returned object (attribute ``_keep``)."""
    b = BeamDesc()
    b.element = kw.get("element", _lib.BEAM_DELTA)
    b.array_mode = kw.get("array_mode", _lib.ARRAY_NONE)
    b.dipole_mode = kw.get("dipole_mode", _lib.DIPOLE_GENERAL)
    b.achromatic = int(kw.get("achromatic", 0))
    b.size = float(kw.get("size", 0.0))
    b.pointing = (C.c_double * 3)(*[float(v) for v in kw.get("pointing", (0.0, 0.0, 1.0))])
    b.orientation = (C.c_double * 3)(*[float(v) for v in kw.get("orientation", (1.0, 0.0, 0.0))])
    b.groundplane = float(kw.get("groundplane", 0.0) or 0.0)
    b.ground_scale = float(kw.get("ground_scale", 0.0) or 0.0)
    b.ground_max = float(kw.get("ground_max", 0.0) or 0.0)
    b.ref_freq_hz = float(kw.get("ref_freq_hz", 0.0))
    b.nax1 = int(kw.get("nax1", 0)); b.nax2 = int(kw.get("nax2", 0))
    b.sep1 = float(kw.get("sep1", 0.0)); b.sep2 = float(kw.get("sep2", 0.0))
    b.east2ax1_deg = float(kw.get("east2ax1_deg", 0.0))
    b.array_pointing = (C.c_double * 3)(*[float(v) for v in kw.get("array_pointing", (0.0, 0.0, 1.0))])
    b.n_elements = int(kw.get("n_elements", 0)); b.nrand = int(kw.get("nrand", 1))
    keep = []
    for name in ("d_element_locs", "d_delays", "d_gains", "d_logmax"):
        t = kw.get(name, None)
        keep.append(t)
        setattr(b, name, None if t is None else t.data_ptr())
    b._keep = keep
    return b


def amp_table(dircos, index, nsrc, spectrum, beam, freqs_hz, pbeam=None, device=None, dtype=torch.float32):
    """``pb200_amp_table``: fp32 amplitude table (slab layout) for the culled sources.
    `spectrum` is a dict of device tensors {flux_scale, index, freq_ref[, flux_offset]} indexed by
    catalogue index, or {spectrum: [nsrc0,nchan]}.  Replaces interferometry.py:6249-6254."""
    device = _dev(device)
    ctx = get_context(device)
    freqs = _h64(freqs_hz)
    nchan = freqs.size
    nbytes = ctx.lib.pb200_amp_bytes(int(nsrc), int(nchan))
    amp = torch.empty((nbytes // 4,), dtype=dtype, device="cuda:{0}".format(device))
    amp_dtype = _lib.AMP_F64 if dtype == torch.float64 else _lib.AMP_F32
    sd = SpectrumDesc()
    for key, field in (("flux_scale", "d_flux_scale"), ("index", "d_index"), ("freq_ref", "d_freq_ref"),
                       ("flux_offset", "d_flux_offset"), ("spectrum", "d_spectrum")):
        t = spectrum.get(key, None)
        setattr(sd, field, None if t is None else t.data_ptr())
    ctx.check(ctx.lib.pb200_amp_table(ctx.handle, _ptr(dircos), _ptr(index), int(nsrc), C.byref(sd), C.byref(beam),
                                      _ptr(pbeam), _ptr(freqs), int(nchan), amp_dtype, _ptr(amp), ctx.stream()))
    return amp


def amp_scale(amp, nsrc, nchan, dircos, component, out=None, device=None):
    """``pb200_amp_scale``: the amplitude table with source row s multiplied by dircos[s, component] -- the table of
    one component of the visibility gradient w.r.t. the baseline vector (interferometry.py:6343)."""
    device = _dev(device)
    ctx = get_context(device)
    if out is None:
        out = torch.empty_like(amp)
    amp_dtype = _lib.AMP_F64 if amp.dtype == torch.float64 else _lib.AMP_F32
    col = C.c_void_p(dircos.data_ptr() + 8 * int(component))
    ctx.check(ctx.lib.pb200_amp_scale(ctx.handle, _ptr(amp), amp_dtype, int(nsrc), int(nchan), col, 3, _ptr(out), ctx.stream()))
    return out


def amp_table_to_dense(amp, nsrc, nchan):
    """Undo the slab layout: returns a [nsrc, nchan] tensor of the table's dtype (testing / inspection)."""
    nsrc_pad = ((max(nsrc, 1) + _lib.SRC_TILE - 1) // _lib.SRC_TILE) * _lib.SRC_TILE
    nslab = (nchan + _lib.SLAB - 1) // _lib.SLAB
    a = amp.view(nslab, nsrc_pad, _lib.SLAB).permute(1, 0, 2).reshape(nsrc_pad, nslab * _lib.SLAB)
    return a[:nsrc, :nchan].contiguous()


def dense_to_amp_table(dense, dtype=torch.float32):
    """[nsrc, nchan] (any float dtype, CUDA) -> slab layout table of `dtype`."""
    nsrc, nchan = dense.shape
    nsrc_pad = ((max(nsrc, 1) + _lib.SRC_TILE - 1) // _lib.SRC_TILE) * _lib.SRC_TILE
    nslab = (nchan + _lib.SLAB - 1) // _lib.SLAB
    full = torch.zeros((nsrc_pad, nslab * _lib.SLAB), dtype=dtype, device=dense.device)
    full[:nsrc, :nchan] = dense.to(dtype)
    return full.view(nsrc_pad, nslab, _lib.SLAB).permute(1, 0, 2).contiguous().view(-1)


def skyvis(dircos, amp, nsrc, baselines_enu, pc_dircos, freqs_hz, src_fwhm_deg=None, method="auto", out=None,
           device=None, nsrc_bright=0):
    """``pb200_skyvis``: V[nbl,nchan] complex128.  Replaces interferometry.py:6155-6165, :6255,
    :6258-6283, :6332-6340.  nsrc_bright: the first nsrc_bright sources are the brightest (see ``brightness_order``)."""
    device = _dev(device)
    ctx = get_context(device)
    bl = _f64(baselines_enu, device)
    freqs = _h64(freqs_hz)
    pc = _h64(pc_dircos)
    nbl, nchan = bl.shape[0], freqs.size
    if out is None:
        out = torch.empty((nbl, nchan), dtype=torch.complex128, device=bl.device)
    if tuple(out.shape) != (nbl, nchan) or out.dtype != torch.complex128 or (nchan > 1 and out.stride(1) != 1) or \
            (nbl > 1 and out.stride(0) < nchan):
        raise ValueError("out must be a [nbl, nchan] complex128 CUDA tensor with unit channel stride (rows may be strided)")
    row_stride = out.stride(0) if nbl > 1 else nchan
    code = {"auto": _lib.SKYVIS_AUTO, "recurrence": _lib.SKYVIS_RECURRENCE, "direct": _lib.SKYVIS_DIRECT,
            "recurrence_scalar": _lib.SKYVIS_RECURRENCE_SCALAR, "fp64": _lib.SKYVIS_FP64,
            "recurrence_lift": _lib.SKYVIS_RECURRENCE_LIFT, "recurrence_3term": _lib.SKYVIS_RECURRENCE_3TERM, "recurrence_3term_scalar": _lib.SKYVIS_RECURRENCE_3TERM_SCALAR,
            "recurrence_quarter": _lib.SKYVIS_RECURRENCE_QUARTER, "recurrence_pair": _lib.SKYVIS_RECURRENCE_PAIR}[method]
    amp_dtype = _lib.AMP_F64 if (amp is not None and amp.dtype == torch.float64) else _lib.AMP_F32
    ctx.check(ctx.lib.pb200_skyvis(ctx.handle, _ptr(dircos), _ptr(amp), amp_dtype, int(nsrc), _ptr(bl), int(nbl), _ptr(pc),
                                   _ptr(freqs), int(nchan), _ptr(src_fwhm_deg), int(nsrc_bright), _ptr(out), int(row_stride), code, ctx.stream()))
    return out


def brightness_order(dircos, index, nsrc, spectrum, beam, freqs_hz, pbeam=None, device=None, power_fraction=0.97, max_fraction=0.05):
    """Order in which the phase sum should take the culled sources: descending amplitude (flux x beam at the centre
    channel, from a one-channel ``pb200_amp_table``).  Returns (perm [nsrc] int64 CUDA tensor, nsrc_bright): the first
    nsrc_bright sources of the permuted list carry `power_fraction` of sum a^2 (at most `max_fraction` of the sources);
    ``pb200_skyvis`` flushes their fp32 partial sums after every tile.  On a GLEAM-like catalogue 1 % of the sources
    hold ~90 % of the power; on a smooth diffuse sky the cap applies and the ordering is harmless."""
    device = _dev(device)
    freqs = _h64(freqs_hz)
    mid = freqs[freqs.size // 2: freqs.size // 2 + 1]
    pb_mid = None if pbeam is None else pbeam[:, freqs.size // 2: freqs.size // 2 + 1].contiguous()
    col = amp_table(dircos, index, nsrc, spectrum, beam, mid, pbeam=pb_mid, device=device)
    key = col.view(-1, _lib.SLAB)[:nsrc, 0].double().square()
    order = torch.argsort(key, descending=True, stable=True)
    csum = torch.cumsum(key[order], dim=0)
    total = csum[-1]
    nb = int(torch.searchsorted(csum, power_fraction * total).item()) + 1 if float(total) > 0.0 else 0
    return order, min(nb, int(max_fraction * nsrc))


def _bcast_strides(t, nbl, nchan):
    """(row, col) element strides of a tensor broadcastable to [nbl, nchan]."""
    if t is None or t.numel() == 1:
        return 0, 0
    if t.ndim == 1:
        if t.numel() == nchan:
            return 0, 1
        if t.numel() == nbl:
            return 1, 0
        raise ValueError("1-D array is neither [nchan] nor [nbl]")
    if tuple(t.shape) == (nbl, nchan):
        return nchan, 1
    if tuple(t.shape) == (1, nchan):
        return 0, 1
    if tuple(t.shape) == (nbl, 1):
        return 1, 0
    raise ValueError("array of shape {0} does not broadcast to [{1},{2}]".format(tuple(t.shape), nbl, nchan))


def noise(skyvis_t, tsys, aeff, effq, df, t_acc, seed, nbl, nchan, snapshot=0, bl_offset=0, nbl_total=None, gains=None,
          flux_unit_k=False, want=("rms", "noise", "vis"), device=None, bl_step=1, out=None):
    """``pb200_noise`` for one snapshot.  tsys / aeff / effq are contiguous fp64 CUDA tensors of
    shape [nbl,nchan], [nchan], [nbl] or scalar (broadcast through strides).
    Replaces interferometry.py:6676-6693 and :6707-6722."""
    device = tsys.device.index if device is None else _dev(device)
    ctx = get_context(device)
    nbl_total = nbl if nbl_total is None else int(nbl_total)
    dev = "cuda:{0}".format(device)
    out = out or {}                                  # optional preallocated outputs {'rms', 'noise', 'vis'} (contiguous, right shape)

    def product(name, dtype):
        if name not in want:
            return None
        t = out.get(name, None)
        if t is None:
            return torch.empty((nbl, nchan), dtype=dtype, device=dev)
        if tuple(t.shape) != (nbl, nchan) or t.dtype != dtype or not t.is_contiguous():
            raise ValueError("preallocated '{0}' must be a contiguous [nbl, nchan] {1} CUDA tensor".format(name, dtype))
        return t

    rms, nz, vis = product("rms", torch.float64), product("noise", torch.complex128), product("vis", torch.complex128)
    st = []
    for t in (tsys, aeff, effq):
        st.extend(_bcast_strides(t, nbl, nchan))
    strides = (C.c_longlong * 6)(*st)
    ctx.check(ctx.lib.pb200_noise(ctx.handle, _ptr(skyvis_t), _ptr(tsys), _ptr(aeff), _ptr(effq), strides, _ptr(gains),
                                  int(nbl), int(nchan), float(df), float(t_acc), int(bool(flux_unit_k)),
                                  int(seed) & 0xFFFFFFFFFFFFFFFF, int(snapshot), int(bl_offset), int(bl_step), nbl_total, 0,
                                  _ptr(rms), _ptr(nz), _ptr(vis), ctx.stream()))
    return rms, nz, vis


def add_noise(skyvis_t, noise_t, gains=None, out=None):
    """``pb200_noise`` in add-only mode: vis = gains*skyvis + noise (interferometry.py:6722)."""
    device = skyvis_t.device.index
    ctx = get_context(device)
    nbl, nchan = skyvis_t.shape
    vis = torch.empty_like(skyvis_t) if out is None else out
    ctx.check(ctx.lib.pb200_noise(ctx.handle, _ptr(skyvis_t), None, None, None, None, _ptr(gains), int(nbl), int(nchan),
                                  1.0, 1.0, 0, 0, 0, 0, 1, nbl, 1, None, _ptr(noise_t), _ptr(vis), ctx.stream()))
    return vis


def delay_nout(nchan, pad=1.0, downsample=True):
    return int(_lib.load().pb200_delay_nout(int(nchan), float(pad), int(bool(downsample))))


def delay_transform(x, bp, wts, df, pad=1.0, downsample=True, nrows=None, nchan=None, device=None, out=None):
    """``pb200_delay_transform``: x [nrows,nchan] complex128 or None; bp / wts [nrows,nchan] or
    [nchan] (broadcast) float64 or None.  Returns [nrows, nout] complex128.
    Replaces interferometry.py:8114-8134."""
    ref = x if x is not None else (bp if bp is not None else wts)
    device = ref.device.index if device is None else _dev(device)
    ctx = get_context(device)
    if x is not None:
        nrows, nchan = x.shape

    def stride(t):
        if t is None:
            return 0
        return 0 if t.ndim == 1 or t.shape[0] == 1 else nchan

    nout = delay_nout(nchan, pad, downsample)
    if out is None:
        out = torch.empty((nrows, nout), dtype=torch.complex128, device="cuda:{0}".format(device))
    elif tuple(out.shape) != (nrows, nout) or out.dtype != torch.complex128 or not out.is_contiguous():
        raise ValueError("out must be a contiguous [nrows, nout] complex128 CUDA tensor")
    ctx.check(ctx.lib.pb200_delay_transform(ctx.handle, _ptr(x), _ptr(bp), stride(bp), _ptr(wts), stride(wts),
                                            int(nrows), int(nchan), float(df), float(pad), int(bool(downsample)),
                                            _ptr(out), ctx.stream()))
    return out


def healpix_beam(map_dev, nside, dircos, nsrc, nchan):
    """``pb200_healpix_beam``: bilinear HEALPix interpolation of a log10 beam map [npix, nchan] (fp32/fp64 CUDA
    tensor, already on the observing channels) at the culled directions.  Returns (logbeam [nsrc,nchan] f64,
    colmax [nchan] f64).  Replaces scripts/run_prisim.py:1897-1905."""
    device = map_dev.device.index
    ctx = get_context(device)
    logbeam = torch.empty((max(nsrc, 1), nchan), dtype=torch.float64, device=map_dev.device)
    colmax = torch.empty((nchan,), dtype=torch.float64, device=map_dev.device)
    dt = _lib.AMP_F64 if map_dev.dtype == torch.float64 else _lib.AMP_F32
    ctx.check(ctx.lib.pb200_healpix_beam(ctx.handle, _ptr(map_dev), dt, int(nside), _ptr(dircos), int(nsrc), int(nchan),
                                         _ptr(logbeam), _ptr(colmax), ctx.stream()))
    return logbeam[:nsrc], colmax


def phase_rotate(vis, baselines_dev, dpos_dircos, freqs_hz):
    """``pb200_phase_rotate``: in-place V[b,f] *= exp(-2 pi i f b.dpos/c) (interferometry.py:7869-7881)."""
    device = vis.device.index
    ctx = get_context(device)
    nbl, nchan = vis.shape
    dpos, freqs = _h64(dpos_dircos), _h64(freqs_hz)      # keep the converted host arrays alive across the call
    ctx.check(ctx.lib.pb200_phase_rotate(ctx.handle, _ptr(vis), _ptr(baselines_dev), int(nbl), _ptr(dpos),
                                         _ptr(freqs), int(nchan), ctx.stream()))
    return vis


def channels_uniform(freqs_hz):
    """``pb200_channels_uniform``: the library's own test for a uniformly spaced channel grid (the recurrence and
    fp64 kernels need one)."""
    freqs = _h64(freqs_hz)
    return bool(_lib.load().pb200_channels_uniform(_ptr(freqs), int(freqs.size)))


def microbench(device=None):
    """``pb200_microbench``: measured FP32-FMA / MUFU / FP64 issue rates on this GPU."""
    device = _dev(device)
    ctx = get_context(device)
    out = (C.c_double * 5)()
    ctx.check(ctx.lib.pb200_microbench(ctx.handle, out, 5))
    return {"ffma_per_s": out[0], "ffma_lanes_per_clk_per_sm": out[1], "mufu_per_s": out[2], "dfma_per_s": out[3],
            "sm_clock_hz": out[4], "fp32_tflops": 2.0 * out[0] / 1e12}
