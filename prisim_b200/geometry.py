"""Host-side coordinate helpers for O(1)-sized quantities (pointing centres, orientations).

The per-source transforms run on the GPU (``pb200_sky_cull``); these numpy versions only convert
the handful of directions the host shim needs to describe a snapshot.  Semantics follow the
un-vendored ``astroutils.geometry`` functions the reference calls (interferometry.py:6122-6165):
az from North through East, (l, m, n) = East, North, Up, hour angle positive to the West.
"""
from __future__ import annotations

import numpy as NP


def altaz2dircos(altaz, units="degrees"):
    altaz = NP.asarray(altaz, dtype=NP.float64).reshape(-1, 2)
    if units == "degrees":
        altaz = NP.radians(altaz)
    alt, az = altaz[:, 0], altaz[:, 1]
    return NP.stack((NP.cos(alt) * NP.sin(az), NP.cos(alt) * NP.cos(az), NP.sin(alt)), axis=1)


def dircos2altaz(dircos, units="degrees"):
    dircos = NP.asarray(dircos, dtype=NP.float64).reshape(-1, 3)
    alt = NP.arcsin(NP.clip(dircos[:, 2], -1.0, 1.0))
    az = NP.arctan2(dircos[:, 0], dircos[:, 1]) % (2 * NP.pi)
    out = NP.stack((alt, az), axis=1)
    return NP.degrees(out) if units == "degrees" else out


def hadec2altaz(hadec, latitude, units="degrees"):
    hadec = NP.asarray(hadec, dtype=NP.float64).reshape(-1, 2)
    if units == "degrees":
        ha, dec, lat = NP.radians(hadec[:, 0]), NP.radians(hadec[:, 1]), NP.radians(latitude)
    else:
        ha, dec, lat = hadec[:, 0], hadec[:, 1], latitude
    north = NP.sin(dec) * NP.cos(lat) - NP.cos(dec) * NP.cos(ha) * NP.sin(lat)
    east = -NP.cos(dec) * NP.sin(ha)
    up = NP.sin(dec) * NP.sin(lat) + NP.cos(dec) * NP.cos(ha) * NP.cos(lat)
    out = NP.stack((NP.arcsin(NP.clip(up, -1.0, 1.0)), NP.arctan2(east, north) % (2 * NP.pi)), axis=1)
    return NP.degrees(out) if units == "degrees" else out


def altaz2hadec(altaz, latitude, units="degrees"):
    altaz = NP.asarray(altaz, dtype=NP.float64).reshape(-1, 2)
    if units == "degrees":
        alt, az, lat = NP.radians(altaz[:, 0]), NP.radians(altaz[:, 1]), NP.radians(latitude)
    else:
        alt, az, lat = altaz[:, 0], altaz[:, 1], latitude
    e, n, u = NP.cos(alt) * NP.sin(az), NP.cos(alt) * NP.cos(az), NP.sin(alt)
    z = n * NP.cos(lat) + u * NP.sin(lat)
    x = -n * NP.sin(lat) + u * NP.cos(lat)
    y = -e
    out = NP.stack((NP.arctan2(y, x) % (2 * NP.pi), NP.arcsin(NP.clip(z, -1.0, 1.0))), axis=1)
    return NP.degrees(out) if units == "degrees" else out


def xyz2enu(xyz, latitude, units="degrees"):
    """Equatorial XYZ (X to HA=0/Dec=0, Y East, Z pole) -> local ENU (interferometry.py:6153)."""
    xyz = NP.asarray(xyz, dtype=NP.float64).reshape(-1, 3)
    lat = NP.radians(latitude) if units == "degrees" else latitude
    return NP.stack((xyz[:, 1], -NP.sin(lat) * xyz[:, 0] + NP.cos(lat) * xyz[:, 2],
                     NP.cos(lat) * xyz[:, 0] + NP.sin(lat) * xyz[:, 2]), axis=1)


def enu2xyz(enu, latitude, units="degrees"):
    enu = NP.asarray(enu, dtype=NP.float64).reshape(-1, 3)
    lat = NP.radians(latitude) if units == "degrees" else latitude
    return NP.stack((-NP.sin(lat) * enu[:, 1] + NP.cos(lat) * enu[:, 2], enu[:, 0],
                     NP.cos(lat) * enu[:, 1] + NP.sin(lat) * enu[:, 2]), axis=1)
