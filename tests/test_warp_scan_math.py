"""The upward sweep of adjust_saturation_profile! (soil_hydrology.jl:192-199) as a scan -- the algebra the warp-per-column kernel
relies on in fast math (csrc/warp_kernel.cuh), checked on the CPU in numpy.

Sequential form (reference): the excess of layer k, e_k = max(s_k + c_{k-1} - 1, 0), is handed to the layer above scaled by the
thickness ratio, c_k = r_k e_k. The carry out of a layer as a function of the carry in is h_k(c) = max(r_k c + r_k (s_k - 1), 0),
a member of the family max(a c + b, g), which is closed under composition:
(a2, b2, g2) o (a1, b1, g1) = (a2 a1, a2 b1 + b2, max(a2 g1 + b2, g2)). An inclusive scan of the h_k evaluated at c = 0 gives every
carry in log2(32) rounds."""
import numpy as np


def sequential(s, dz):
    s = s.copy()
    carry = 0.0
    carries = np.zeros(len(s))
    for k in range(len(s)):
        s[k] += carry
        if k < len(s) - 1:
            e = max(s[k] - 1.0, 0.0)
            s[k] -= e
            carry = e * dz[k] / dz[k + 1]
            carries[k] = carry
    return s, carries


def scan(s, dz):
    n = len(s)
    lanes = 32
    al, be, ga = np.zeros(lanes), np.zeros(lanes), np.zeros(lanes)
    for k in range(n - 1):                      # every layer but the top one hands its excess on
        r = dz[k] / dz[k + 1]
        al[k], be[k] = r, r * (s[k] - 1.0)
    d = 1
    while d < lanes:                            # Hillis-Steele inclusive scan, lane l combines with lane l - d
        a1, b1, g1 = np.roll(al, d), np.roll(be, d), np.roll(ga, d)
        act = np.arange(lanes) >= d
        ga = np.where(act, np.maximum(al * g1 + be, ga), ga)
        be = np.where(act, al * b1 + be, be)
        al = np.where(act, al * a1, al)
        d *= 2
    cout = np.maximum(be, ga)                   # H_l(0): carry out of layer l
    cin = np.concatenate([[0.0], cout[:-1]])[:n]
    out = s + cin
    e = np.maximum(out[:-1] - 1.0, 0.0)
    out[:-1] -= e
    return out, cout[:n]


def test_scan_equals_the_sequential_sweep():
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(1, 32))
        dz = np.cumsum(rng.uniform(0.05, 2.0, n + 1))[:n] if trial % 2 else rng.uniform(0.05, 3.0, n)
        s = rng.uniform(0.0, 1.0, n)
        hot = rng.random(n) < 0.5
        s[hot] = 1.0 + rng.uniform(-1e-3, 0.2, hot.sum())      # saturated zones with inflow, chains of any length
        if trial % 5 == 0:
            s[:] = 1.0 + rng.uniform(0.0, 0.05, n)              # the whole column over-saturated
        a, ca = sequential(s, dz)
        b, cb = scan(s, dz)
        assert np.allclose(a, b, rtol=0, atol=1e-13), (trial, np.max(np.abs(a - b)))
        assert np.allclose(ca[: n - 1], cb[: n - 1], rtol=1e-12, atol=1e-15)
        assert np.all(b[:-1] <= 1.0)                             # every layer but the top one ends at or below saturation
        # water handed on is conserved: sum s dz before == after (the top layer keeps its excess for surface_excess_water)
        assert abs(np.dot(s, dz) - np.dot(b, dz)) <= 1e-12 * np.dot(s, dz)


def test_composition_rule():
    rng = np.random.default_rng(3)
    for _ in range(100):
        a1, a2 = rng.uniform(0.1, 3.0, 2)
        b1, b2 = rng.uniform(-2.0, 2.0, 2)
        g1, g2 = rng.uniform(0.0, 1.0, 2)
        c = rng.uniform(-1.0, 3.0, 16)
        lhs = np.maximum(a2 * np.maximum(a1 * c + b1, g1) + b2, g2)
        rhs = np.maximum((a2 * a1) * c + (a2 * b1 + b2), max(a2 * g1 + b2, g2))
        assert np.allclose(lhs, rhs, rtol=1e-14, atol=1e-14)
