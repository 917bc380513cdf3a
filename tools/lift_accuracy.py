"""fp32 emulation (numpy, CPU) of the two-channel phasor step of the phase-sum kernel: the 4-op complex rotation by r^2
against the 3-op lifted rotation (shears x += t y; y += s x; x += t y, t = -tan(phi/2), s = sin(phi)), as a function of the
largest step angle.  Prints rms and max |error| of the phasor over a 32-channel block (16 steps) for each scheme; this is
where PB_LIFT_MAX_ANGLE = 2.0 rad in csrc/skyvis.cu comes from."""
import numpy as np
f32 = np.float32
rng = np.random.default_rng(1)


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def run(theta2max, n=200000):
    phi1 = rng.uniform(-theta2max / 2, theta2max / 2, n)            # per-channel angle
    ph0 = rng.uniform(-np.pi, np.pi, n)
    c1 = np.cos(phi1).astype(f32); s1 = np.sin(phi1).astype(f32)
    x = np.stack([np.cos(ph0), np.cos(ph0 + phi1)], 1).astype(f32); y = np.stack([np.sin(ph0), np.sin(ph0 + phi1)], 1).astype(f32)
    one = np.ones(2, f32)
    t = (-(s1 / c1)).astype(f32)[:, None] * one; s = (f32(2) * s1 * c1).astype(f32)[:, None] * one
    r2r = (c1 * c1 - s1 * s1).astype(f32)[:, None] * one; r2i = (f32(2) * c1 * s1).astype(f32)[:, None] * one
    xl, yl, xc, yc = x.copy(), y.copy(), x.copy(), y.copy()
    out = [0.0] * 4
    for j in range(1, 16):
        x1 = fma(t, yl, xl); yl = fma(s, x1, yl); xl = fma(t, yl, x1)
        t1 = (xc * r2r).astype(f32); t2 = (xc * r2i).astype(f32)
        xc, yc = fma(-yc, r2i, t1), fma(yc, r2r, t2)
        k = np.array([2 * j, 2 * j + 1])
        ex = np.cos(ph0[:, None] + k[None, :] * phi1[:, None]); ey = np.sin(ph0[:, None] + k[None, :] * phi1[:, None])
        el, ec = np.hypot(xl - ex, yl - ey), np.hypot(xc - ex, yc - ey)
        out = [max(out[0], np.sqrt(np.mean(el ** 2))), max(out[1], np.sqrt(np.mean(ec ** 2))), max(out[2], el.max()), max(out[3], ec.max())]
    return out


print("max step angle [rad] | rms lifted, rms 4-op | max lifted, max 4-op")
for th in (0.5, 1.0, 1.57, 2.0, 2.5, 2.8, 3.0):
    print(th, " ".join("%.2e" % v for v in run(th)))
