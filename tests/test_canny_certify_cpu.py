"""CPU check of the certification in canny_grad_nms_spec2_kernel (image_b200/csrc/canny.cu).

Tier 1 of that kernel decides a pixel's class in fp32 only when every comparison clears a tolerance; this
test re-states tier 1 in numpy float32 (same operation order, fused multiply-adds emulated through float64),
perturbs the two approximate instructions (rcp.approx, sqrt.approx) to both ends of their error interval,
and asserts on adversarial planes that a DECIDED class always equals the exact class of the oracle
(orc_canny_gradient + orc_canny_maxima, the reference's double arithmetic with the libm hypot / atan2).
Undecided pixels go to the exact tier 2 in the kernel, so only wrong certainties could break parity."""
import ctypes as C

import numpy as np
import pytest

f32 = np.float32


def fma32(a, b, c):
    """fl32(a*b + c) with one rounding (the product of two float32 is exact in float64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def exact_classes(po, D, acc, low, high):
    lib = po.lib("oracle")
    ny, nx = D.shape
    grad = np.zeros((ny, nx), np.float64); theta = np.zeros((ny, nx), np.float64)
    cls = np.zeros((ny, nx), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.orc_canny_gradient(p(D), nx, ny, int(acc), p(grad), p(theta))
    lib.orc_canny_maxima(p(grad), p(theta), nx, ny, int(low), int(high), p(cls))
    return cls


def tier1(D, acc, low, high, d_sqrt, d_rcp):
    """Returns (cls, decided) for the interior of plane D (float32), tiles of 32x32 with origin (2, 2).
    d_sqrt / d_rcp: relative perturbation applied to the approximate sqrt / reciprocal."""
    ny, nx = D.shape
    two = f32(2.0)
    # gradient everywhere it is defined without clamping
    c = D[1:-1, 1:-1]
    l, r = D[1:-1, :-2], D[1:-1, 2:]
    ul, um, ur = D[:-2, :-2], D[:-2, 1:-1], D[:-2, 2:]
    dl, dm, dr = D[2:, :-2], D[2:, 1:-1], D[2:, 2:]
    dh1 = r - l
    vy = dm - um
    if acc:
        dh0, dh2 = ur - ul, dr - dl
        vp, vm = dr - ur, dl - ul
        h = fma32(np.full_like(dh1, two), dh1, dh2 + dh0)
        v = fma32(np.full_like(vy, two), vy, vp + vm)
        S = fma32(np.full_like(vy, two), np.abs(dh1) + np.abs(vy), (np.abs(dh2) + np.abs(dh0)) + (np.abs(vp) + np.abs(vm)))
    else:
        h, v = dh1, vy
        S = np.abs(h) + np.abs(v)
    q = fma32(h, h, v * v)
    g = (np.sqrt(q.astype(np.float64)) * (1.0 + d_sqrt)).astype(f32)
    e = fma32(np.full_like(S, f32(6e-7)), S, f32(4e-7) * g)
    G = np.zeros((ny, nx), f32); H = np.zeros((ny, nx), f32); V = np.zeros((ny, nx), f32); Ee = np.zeros((ny, nx), f32)
    G[1:-1, 1:-1], H[1:-1, 1:-1], V[1:-1, 1:-1], Ee[1:-1, 1:-1] = g, h, v, e
    cls = np.zeros((ny, nx), np.uint8)
    decided = np.zeros((ny, nx), bool)
    lowf, highf = f32(low), f32(high)
    for y0 in range(2, ny - 2 - 32 + 1, 32):
        for x0 in range(2, nx - 2 - 32 + 1, 32):
            E = max(f32(Ee[y0 - 1:y0 + 33, x0 - 1:x0 + 33].max()), f32(1e-7))     # the 34 x 34 gradient tile
            T0 = f32(2.0) * E
            ys, xs = np.mgrid[y0:y0 + 32, x0:x0 + 32]
            now = G[ys, xs]
            live = now >= lowf - T0
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = (1.0 / now.astype(np.float64) * (1.0 + d_rcp)).astype(f32)
            inv = np.where(np.isfinite(inv), inv, f32(0))                  # (now == 0 is never `live` unless low - T0 <= 0; then it is ambiguous)
            cs, sn = H[ys, xs] * inv, V[ys, xs] * inv
            dcs = f32(8.0) * fma32(np.full_like(inv, E), inv, np.full_like(inv, f32(5e-7)))
            ngx, ngy = (cs < 0).astype(int), (sn < 0).astype(int)
            wbx = cs + ngx.astype(f32); wax = f32(1.0) - wbx
            wby = sn + ngy.astype(f32); way = f32(1.0) - wby
            py, px = ys - ngy, xs - ngx
            my, mx = ys + ngy - 1, xs + ngx - 1
            p11, p12, p21, p22 = G[py, px], G[py, px + 1], G[py + 1, px], G[py + 1, px + 1]
            m11, m12, m21, m22 = G[my, mx], G[my, mx + 1], G[my + 1, mx], G[my + 1, mx + 1]
            nbp = way * fma32(wax, p11, wbx * p12) + wby * fma32(wax, p21, wbx * p22)
            nbm = wby * fma32(wbx, m11, wax * m12) + way * fma32(wbx, m21, wax * m22)
            spp = np.maximum(np.maximum(p11, p12), np.maximum(p21, p22)) - np.minimum(np.minimum(p11, p12), np.minimum(p21, p22))
            spm = np.maximum(np.maximum(m11, m12), np.maximum(m21, m22)) - np.minimum(np.minimum(m11, m12), np.minimum(m21, m22))
            tolp = fma32(dcs, spp, np.full_like(spp, f32(2.0) * T0)); tolm = fma32(dcs, spm, np.full_like(spm, f32(2.0) * T0))
            amb = (now <= lowf + T0) | (np.minimum(np.abs(cs), np.abs(sn)) <= dcs)
            sup = now < np.maximum(nbp - tolp, nbm - tolm)
            top = now > np.maximum(nbp + tolp, nbm + tolm)
            hi2, hi1 = now >= highf + T0, now < highf - T0
            und = amb | (~sup & (~top | (~hi2 & ~hi1)))
            c1 = np.where(~und & ~sup, np.where(hi2, 2, 1), 0).astype(np.uint8)
            cls[ys, xs] = np.where(live, c1, 0)
            decided[ys, xs] = np.where(live, ~und, True)
    return cls, decided


def planes(rng):
    n = 2 + 32 * 3 + 2
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    out = []
    noise = rng.random((n + 8, n + 8)) * 16
    k = np.exp(-np.arange(-4, 5) ** 2 / 4.0); k /= k.sum()
    sm = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 0, noise))[4:-4, 4:-4]
    out.append(("blurred noise (like the bench frames)", 96 + sm))
    for slope in (3.0 / 8, 10.0 / 8, 3.0 / 2, 10.0 / 2):            # gradient magnitudes right at the two thresholds
        out.append(("ramp slope %.4f" % slope, 20 + slope * xx + rng.random((n, n)) * 1e-5))
        out.append(("ramp slope %.4f (y)" % slope, 20 + slope * yy + rng.random((n, n)) * 1e-6))
        out.append(("diagonal ramp %.4f" % slope, 20 + slope * (xx + yy) / np.sqrt(2.0)))
    out.append(("plateau with 1e-6 ripples", 100 + rng.random((n, n)) * 1e-6))
    out.append(("constant", np.full((n, n), 77.0)))
    out.append(("checkerboard", 100 + 20.0 * ((xx.astype(int) + yy.astype(int)) % 2)))
    out.append(("ridge with equal neighbours", 50 + 30.0 * np.exp(-((xx - n / 2) ** 2) / 18.0)))
    out.append(("steps", 40 + 25.0 * (xx.astype(int) // 7 % 2) + 13.0 * (yy.astype(int) // 5 % 2)))
    out.append(("sinusoids", 128 + 60 * np.sin(xx / 3.1) * np.cos(yy / 4.7) + sm * 0.05))
    return [(name, p.astype(f32)) for name, p in out]


@pytest.mark.parametrize("acc", [True, False])
def test_certified_decisions_equal_the_exact_classes(oracle, acc):
    rng = np.random.default_rng(11)
    total = und = 0
    for name, D in planes(rng):
        D = np.ascontiguousarray(D)
        for low, high in ((3, 10), (1, 2), (0, 40)):
            ex = exact_classes(oracle, D, acc, low, high)
            for d_sqrt in (-2.0 ** -23, 0.0, 2.0 ** -23):
                for d_rcp in (-2.0 ** -23, 2.0 ** -23):
                    cls, dec = tier1(D, acc, low, high, d_sqrt, d_rcp)
                    core = np.zeros_like(dec); core[2:98, 2:98] = True
                    bad = core & dec & (cls != ex)
                    assert not bad.any(), "%s (low %d, high %d): %d certified pixels differ from the exact class, first at %s" % (
                        name, low, high, int(bad.sum()), tuple(np.argwhere(bad)[0]))
            total += 96 * 96; und += int((core & ~dec).sum())
    assert und < total            # (tier 1 does decide something: the test is not vacuous)


def test_the_check_has_teeth(oracle):
    """With the tolerance constants shrunk by 1e5 the same emulation must produce wrong certainties on the
    adversarial planes -- otherwise the test above would prove nothing."""
    import re
    src = open(__file__).read().split("@pytest.mark.parametrize")[0]
    weak = re.sub(r"f32\((6e-7|4e-7|5e-7|1e-7)\)", lambda m: "f32(%s)" % m.group(1).replace("e-7", "e-12"), src)
    ns = {}
    exec(compile(weak, "weakened", "exec"), ns)
    rng = np.random.default_rng(11)
    wrong = 0
    for name, D in ns["planes"](rng):
        D = np.ascontiguousarray(D)
        ex = ns["exact_classes"](oracle, D, True, 3, 10)
        cls, dec = ns["tier1"](D, True, 3, 10, 2.0 ** -23, -2.0 ** -23)
        wrong += int((dec & (cls != ex))[2:98, 2:98].sum())
    assert wrong > 100
