"""Minimal FITS reader / writer (primary HDU + IMAGE extensions) in plain numpy.

PRISim exchanges the per-snapshot region-of-interest tables between ranks through a FITS file written by
``ROI_parameters.save`` with astropy (interferometry.py:4621-4723: a primary header with the telescope keywords and one
``ImageHDU`` per array -- ``FREQ``, ``IND_j``, ``PB_j``, ``DELAYS_j``, ...) and read back with
``fits.getdata(roifile, extname=...)`` (scripts/run_prisim.py:1959-1961, :2181-2182) and by ``ROI_parameters(init_file=)``
(:4080-4205).  astropy is not available in this image, and this interchange needs only the fixed-format subset of the FITS
standard (v4.0): 2880-byte blocks of 80-character header cards, big-endian image data, ``XTENSION = 'IMAGE'`` extensions with
``EXTNAME``, and the ESO ``HIERARCH`` convention astropy uses for keywords longer than eight characters
(``element_shape`` ...).  Everything written here follows the standard's fixed format, so astropy / cfitsio read it; files
written by astropy with image extensions only are read back (no tables, no scaling keywords, no compression).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as NP

BLOCK = 2880
_BITPIX = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}
_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def _fmt_value(value):
    if isinstance(value, (bool, NP.bool_)):
        return "{0:>20s}".format("T" if value else "F")
    if isinstance(value, (int, NP.integer)):
        return "{0:>20d}".format(int(value))
    if isinstance(value, (float, NP.floating)):
        text = repr(float(value)).upper()
        if "E" not in text and "." not in text and "N" not in text:
            text += ".0"
        return "{0:>20s}".format(text)
    text = "'" + str(value).replace("'", "''")
    return "{0:<9s}'".format(text)                                          # opening quote + at least 8 characters + closing quote


def _card(key, value, comment=None):
    key = str(key)
    if len(key) <= 8 and key.upper() == key and " " not in key:
        head = "{0:<8s}= ".format(key)
    elif len(key) <= 8 and " " not in key:
        head = "{0:<8s}= ".format(key.upper())
    else:
        head = "HIERARCH {0} = ".format(key)                               # ESO convention, as astropy does for long keywords
    card = head + _fmt_value(value)
    if comment:
        card += " / " + str(comment)
    if len(card) > 80:
        card = card[:80]
        if card.count("'") % 2:                                              # never cut a string value open
            raise ValueError("header card for {0!r} does not fit in 80 characters".format(key))
    return "{0:<80s}".format(card)


def _header_block(cards):
    text = "".join(cards) + "{0:<80s}".format("END")
    pad = (-len(text)) % BLOCK
    return (text + " " * pad).encode("ascii")


def _data_block(data):
    if data is None:
        return b""
    raw = NP.ascontiguousarray(data).astype(data.dtype.newbyteorder(">"), copy=False).tobytes()
    return raw + b"\0" * ((-len(raw)) % BLOCK)


def _shape_cards(data):
    if data is None:
        return [_card("BITPIX", 8, "array data type"), _card("NAXIS", 0, "number of array dimensions")]
    code = data.dtype.kind + str(data.dtype.itemsize)
    if data.dtype.kind == "b":
        raise TypeError("boolean arrays cannot be written to a FITS image; cast to uint8")
    if code not in _BITPIX:
        raise TypeError("dtype {0} cannot be written to a FITS image".format(data.dtype))
    cards = [_card("BITPIX", _BITPIX[code], "array data type"), _card("NAXIS", data.ndim, "number of array dimensions")]
    for i, n in enumerate(reversed(data.shape)):                           # NAXIS1 is the fastest axis = the last numpy axis
        cards.append(_card("NAXIS{0}".format(i + 1), int(n)))
    return cards


def write(filename, primary_header, extensions, overwrite=False):
    """Write a FITS file: an empty primary HDU carrying `primary_header` ({key: value | (value, comment)}) followed by one
    IMAGE extension per entry of `extensions` = [(extname, ndarray, {header})].  Keys longer than 8 characters are written
    with the HIERARCH convention."""
    if os.path.exists(filename) and not overwrite:
        raise IOError("File {0} already exists; pass overwrite=True".format(filename))
    def user_cards(hdr):
        out = []
        for key, val in (hdr or {}).items():
            value, comment = val if isinstance(val, tuple) else (val, None)
            out.append(_card(key, value, comment))
        return out
    blocks = []
    cards = [_card("SIMPLE", True, "conforms to FITS standard")] + _shape_cards(None) + [_card("EXTEND", True)]
    blocks.append(_header_block(cards + user_cards(primary_header)))
    for name, data, hdr in extensions:
        data = NP.asarray(data)
        if data.dtype == NP.int64 and data.size and NP.abs(data).max() < 2 ** 31:
            pass                                                            # kept as 64-bit: astropy writes what it is given
        cards = [_card("XTENSION", "IMAGE", "Image extension")] + _shape_cards(data)
        cards += [_card("PCOUNT", 0, "number of parameters"), _card("GCOUNT", 1, "number of groups"), _card("EXTNAME", str(name), "extension name")]
        blocks.append(_header_block(cards + user_cards(hdr)))
        blocks.append(_data_block(data))
    with open(filename, "wb") as f:
        for b in blocks:
            f.write(b)


def _parse_value(text):
    text = text.strip()
    if text.startswith("'"):
        end, i = None, 1
        while i < len(text):                                                # closing quote that is not a doubled quote
            if text[i] == "'":
                if i + 1 < len(text) and text[i + 1] == "'":
                    i += 2
                    continue
                end = i
                break
            i += 1
        return text[1:end].replace("''", "'").rstrip()
    text = text.split("/")[0].strip()
    if text == "T":
        return True
    if text == "F":
        return False
    try:
        return int(text)
    except ValueError:
        return float(text.replace("D", "E"))


def _read_header(f):
    cards = OrderedDict()
    while True:
        block = f.read(BLOCK)
        if len(block) < BLOCK:
            return None
        done = False
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode("ascii")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if key == "HIERARCH":
                name, _, rest = card[9:].partition("=")
                cards[name.strip()] = _parse_value(rest)
            elif card[8:10] == "= ":
                cards[key] = _parse_value(card[10:])
        if done:
            return cards


def read(filename):
    """Read every HDU: returns [(header OrderedDict, ndarray | None)]; image data come back in native byte order."""
    hdus = []
    with open(filename, "rb") as f:
        while True:
            hdr = _read_header(f)
            if hdr is None:
                break
            naxis = int(hdr.get("NAXIS", 0))
            shape = tuple(int(hdr["NAXIS{0}".format(i)]) for i in range(naxis, 0, -1))
            data = None
            if naxis > 0:
                dt = NP.dtype(_DTYPE[int(hdr["BITPIX"])])
                count = int(NP.prod(shape))
                nbytes = count * dt.itemsize
                raw = f.read(nbytes + ((-nbytes) % BLOCK))
                data = NP.frombuffer(raw, dtype=dt, count=count).reshape(shape).astype(dt.newbyteorder("="))
            hdus.append((hdr, data))
    return hdus


def getdata(filename, extname):
    """``astropy.io.fits.getdata(filename, extname=...)`` for image extensions (scripts/run_prisim.py:1959-1961)."""
    for hdr, data in read(filename):
        if str(hdr.get("EXTNAME", "")).upper() == str(extname).upper():
            return data
    raise KeyError("Extension {0} not found.".format(extname))


def header_get(hdr, key, default=None):
    """Case-insensitive header lookup (astropy headers are case-insensitive)."""
    for k, v in hdr.items():
        if k.upper() == key.upper():
            return v
    return default
