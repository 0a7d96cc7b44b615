"""Philox4x32-10 and the device sampler's stream, restated in numpy (TEST INFRASTRUCTURE ONLY).

The reference draws its collocation points with torch's CPU generator (train.py:26-39;
poc/main.py:124-156); a device-side sampler cannot reproduce that stream, so its contract is
its own (include/pinn_b200.h, pinn_sample): counter-based Philox4x32-10 (Salmon, Moraes, Dror,
Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11 - the published algorithm, pinned
below by the Random123 known-answer vectors), one counter per point, followed by the
reference's clamp / boundary-set rules restated literally.  Parity of the device sampler with
this file is bit-exact (tests/test_gpu_train.py); parity with the reference is distributional
and semantic (same box, same clamp rule, same sets).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised: uint32 arrays (or scalars) -> 4 uint32 arrays"""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = p0.astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def u01(r):
    """24 random bits -> float32 in [0,1)"""
    return (r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def sample_batch(n, seed, batch, box=(-18, 18, -18, 18, -18, 18, 0.2, 4.0), cutoff=0.005, bcutoff=17.5):
    """The stream of pinn_sample in float32 arithmetic: x, y, z, R, mask (bit0: r1 >= bcutoff, bit1: r2 >= bcutoff),
    counts.  Clamp and sets as in train.py:32-39 / poc/main.py:147-149, 391-393."""
    f = np.float32
    i = np.arange(n, dtype=np.uint64)
    r = philox4x32_10((i & np.uint64(0xFFFFFFFF)).astype(np.uint32), (i >> np.uint64(32)).astype(np.uint32),
                      np.full(n, batch & 0xFFFFFFFF, np.uint32), np.full(n, (batch >> 32) & 0xFFFFFFFF, np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    xL, xR, yL, yR, zL, zR, RL, RR = [f(v) for v in box]

    def fma(a, b, c):  # float32 fused multiply-add (exact product in float64, one rounding)
        return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(np.float32) if np.isscalar(b) or np.ndim(b) == 0 \
            else (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, np.float64)).astype(np.float32)

    x = fma(u01(r[0]), f(xR - xL), xL)
    y = fma(u01(r[1]), f(yR - yL), yL)
    z = fma(u01(r[2]), f(zR - zL), zL)
    R = fma(u01(r[3]), f(RR - RL), RL)
    yz = fma(y, y, (z * z).astype(np.float32))
    c2 = f(f(cutoff) * f(cutoff))
    d1, d2 = (x - R).astype(np.float32), (x + R).astype(np.float32)
    near = (fma(d1, d1, yz) < c2) | (fma(d2, d2, yz) < c2)
    x = np.where(near, f(cutoff), x).astype(np.float32)
    b2 = f(f(bcutoff) * f(bcutoff))
    d1, d2 = (x - R).astype(np.float32), (x + R).astype(np.float32)
    m1 = fma(d1, d1, yz) >= b2
    m2 = fma(d2, d2, yz) >= b2
    mask = (m1.astype(np.uint8) | (m2.astype(np.uint8) << 1)).astype(np.uint8)
    return x, y, z, R, mask, (int(m1.sum()), int(m2.sum()))
