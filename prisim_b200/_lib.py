"""ctypes binding of ``libprisim_b200.so`` (the C-ABI declared in ``include/prisim_b200.h``).

The product path has no CPU fallback: if the shared library is missing, or a compute call is made
without a CUDA device, this module raises.  Device buffers are ``torch`` CUDA tensors; only their
``data_ptr()`` and the current stream handle cross the ABI.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB200_LIB") or os.path.join(_HERE, "libprisim_b200.so")   # PB200_LIB: developer override (tools/variants.sh)

PB200_OK = 0
ABI_VERSION = 200                     # PB200_VERSION of include/prisim_b200.h
SKY_ALTAZ, SKY_HADEC, SKY_DIRCOS = 0, 1, 2
BEAM_DELTA, BEAM_AIRY, BEAM_GAUSSIAN, BEAM_DIPOLE, BEAM_TABLE, BEAM_LOGTABLE = 0, 1, 2, 3, 4, 5
ARRAY_NONE, ARRAY_ANALYTIC, ARRAY_ELEMENTS = 0, 1, 2
DIPOLE_GENERAL, DIPOLE_SHORT, DIPOLE_HALFWAVE = 0, 1, 2
SKYVIS_AUTO, SKYVIS_RECURRENCE, SKYVIS_DIRECT, SKYVIS_RECURRENCE_SCALAR, SKYVIS_FP64, SKYVIS_RECURRENCE_LIFT, SKYVIS_RECURRENCE_3TERM, SKYVIS_RECURRENCE_3TERM_SCALAR = 0, 1, 2, 3, 4, 5, 6, 7
SKYVIS_RECURRENCE_QUARTER = 8
SKYVIS_RECURRENCE_PAIR = 9
SLAB, SRC_TILE = 128, 32
AMP_F32, AMP_F64 = 0, 1


class BeamDesc(C.Structure):
    """Mirror of ``pb200_beam_desc`` (include/prisim_b200.h)."""
    _fields_ = [
        ("element", C.c_int32), ("array_mode", C.c_int32), ("dipole_mode", C.c_int32), ("achromatic", C.c_int32),
        ("size", C.c_double), ("pointing", C.c_double * 3), ("orientation", C.c_double * 3),
        ("groundplane", C.c_double), ("ground_scale", C.c_double), ("ground_max", C.c_double),
        ("ref_freq_hz", C.c_double),
        ("nax1", C.c_int32), ("nax2", C.c_int32),
        ("sep1", C.c_double), ("sep2", C.c_double), ("east2ax1_deg", C.c_double),
        ("array_pointing", C.c_double * 3),
        ("n_elements", C.c_int32), ("nrand", C.c_int32),
        ("d_element_locs", C.c_void_p), ("d_delays", C.c_void_p), ("d_gains", C.c_void_p), ("d_logmax", C.c_void_p),
    ]


class SpectrumDesc(C.Structure):
    """Mirror of ``pb200_spectrum_desc``."""
    _fields_ = [("d_flux_scale", C.c_void_p), ("d_index", C.c_void_p), ("d_freq_ref", C.c_void_p),
                ("d_flux_offset", C.c_void_p), ("d_spectrum", C.c_void_p)]


# every symbol include/prisim_b200.h declares: name -> (restype, argtypes)
_vp, _i, _d, _ll, _u64 = C.c_void_p, C.c_int, C.c_double, C.c_longlong, C.c_uint64
SYMBOLS = {
    "pb200_version": (_i, []),
    "pb200_ctx_create": (_i, [C.POINTER(_vp), _i]),
    "pb200_ctx_destroy": (None, [_vp]),
    "pb200_last_error": (C.c_char_p, [_vp]),
    "pb200_ctx_set_option": (_i, [_vp, C.c_char_p, _ll]),
    "pb200_launch_count": (_ll, [_vp]),
    "pb200_sky_cull": (_i, [_vp, _vp, _i, _i, _d, _d, _vp, _vp, _vp, C.POINTER(_i), _vp]),
    "pb200_amp_bytes": (C.c_size_t, [_i, _i]),
    "pb200_nsrc_pad": (_i, [_i]),
    "pb200_amp_table": (_i, [_vp, _vp, _vp, _i, C.POINTER(SpectrumDesc), C.POINTER(BeamDesc), _vp, _vp, _i, _i, _vp, _vp]),
    "pb200_amp_scale": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "pb200_skyvis": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _ll, _i, _vp]),
    "pb200_channels_uniform": (_i, [_vp, _i]),
    "pb200_noise": (_i, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_ll), _vp, _i, _i, _d, _d, _i, _u64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pb200_delay_nout": (_i, [_i, _d, _i]),
    "pb200_delay_transform": (_i, [_vp, _vp, _vp, _ll, _vp, _ll, _i, _i, _d, _d, _i, _vp, _vp]),
    "pb200_phase_rotate": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "pb200_healpix_beam": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "pb200_device_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "pb200_device_free": (_i, [_vp, _vp]),
    "pb200_peer_export": (_i, [_vp, _vp, _vp]),
    "pb200_peer_open": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "pb200_peer_close": (_i, [_vp, _vp]),
    "pb200_microbench": (_i, [_vp, C.POINTER(_d), _i]),
}

_lib = None


class PB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed for this) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PB200Error(
            "libprisim_b200.so not found at {0}; build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C prisim_b200/csrc`.  There is no CPU fallback.".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)         # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.pb200_version() != ABI_VERSION:
        raise PB200Error("libprisim_b200.so has ABI version {0}, this package needs {1}: rebuild it (make -C prisim_b200/csrc)".format(
            lib.pb200_version(), ABI_VERSION))
    _lib = lib
    return lib


def _ptr(t):
    """Device (or host) address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)       # numpy (host pointers: h_* arguments)


class Context:
    """One ``pb200_ctx`` (one per device per host thread)."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise PB200Error("prisim_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
        self.lib = load()
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.pb200_ctx_create(C.byref(h), self.device)
        if rc != PB200_OK:
            raise PB200Error("pb200_ctx_create(device={0}) failed with code {1}".format(device, rc))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pb200_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != PB200_OK:
            msg = self.lib.pb200_last_error(self.handle)
            raise PB200Error("prisim_b200 error {0}: {1}".format(rc, msg.decode() if msg else ""))

    def set_option(self, name, value):
        """``pb200_ctx_set_option``: developer / test knobs ('skyvis_spc', 'dt_force_r8')."""
        self.check(self.lib.pb200_ctx_set_option(self.handle, name.encode(), int(value)))

    @property
    def launches(self):
        return int(self.lib.pb200_launch_count(self.handle))

    def stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)


_contexts = {}


def get_context(device=0):
    device = int(device)
    ctx = _contexts.get(device)
    if ctx is None or ctx.handle is None:
        ctx = Context(device)
        _contexts[device] = ctx
    return ctx
