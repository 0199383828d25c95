"""Synthetic XYZRGB frames for tests and bench.

Records are PCL's 32-byte ``PointXYZRGB``: x,y,z float32 at 0/4/8, 1.0f at 12, b,g,r,a uint8 at 16..19,
zero padding to 32.  All generators are seeded ``numpy.random.default_rng(seed)`` (PCG64).
"""
import numpy as np

POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"),
                        ("b", "u1"), ("g", "u1"), ("r", "u1"), ("a", "u1"), ("pad", "u1", (12,))])
assert POINT_DTYPE.itemsize == 32

_ELLIPSOIDS = np.array([
    # centre                radii
    [0.50, 0.5, 0.55, 0.12, 0.08, 0.22],
    [0.50, 0.5, 0.85, 0.07, 0.07, 0.08],
    [0.38, 0.5, 0.25, 0.05, 0.05, 0.25],
    [0.62, 0.5, 0.25, 0.05, 0.05, 0.25],
    [0.30, 0.5, 0.60, 0.04, 0.04, 0.20],
    [0.70, 0.5, 0.60, 0.04, 0.04, 0.20]])


def pack_points(xyz, rgb, alpha=255):
    """xyz (n,3) float32, rgb (n,3) uint8 in r,g,b order -> (n,) POINT_DTYPE."""
    n = xyz.shape[0]
    p = np.zeros(n, POINT_DTYPE)
    p["x"], p["y"], p["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    p["w"] = 1.0
    p["r"], p["g"], p["b"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    p["a"] = alpha
    return p


def normalize(xyz, bb_expand_factor=0.2):
    """The reference's normalize_pointclouds for one frame (impl.hpp:1915-1945): bbox expanded by the
    factor on each side, then (x - min) / range in float32."""
    mn = xyz.min(axis=0)
    mx = xyz.max(axis=0)
    ext = (mx - mn).astype(np.float32)
    mn = (mn - np.float32(bb_expand_factor) * ext).astype(np.float32)
    mx = (mx + np.float32(bb_expand_factor) * ext).astype(np.float32)
    rng = (mx - mn).astype(np.float32)
    return ((xyz - mn) / rng).astype(np.float32)


def gen_surface(n, seed=0):
    """G-surf(N, seed): "8iVFB-like" union of six ellipsoid surfaces, smooth colour + noise, shuffled order."""
    rng = np.random.default_rng(seed)
    a, b, c = _ELLIPSOIDS[:, 3], _ELLIPSOIDS[:, 4], _ELLIPSOIDS[:, 5]
    wts = a * b + a * c + b * c
    counts = rng.multinomial(n, wts / wts.sum())
    parts = []
    for e, m in zip(_ELLIPSOIDS, counts):
        d = rng.normal(size=(m, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        parts.append(e[:3] + d * e[3:])
    xyz = np.concatenate(parts).astype(np.float32)
    xyz = normalize(xyz, 0.2)
    x, y, z = xyz[:, 0].astype(np.float64), xyz[:, 1].astype(np.float64), xyz[:, 2].astype(np.float64)
    col = np.stack([128 + 100 * np.sin(40 * x + 9 * z), 128 + 90 * np.cos(31 * y + 17 * z), 128 + 110 * np.sin(23 * z)], axis=1)
    col = np.clip(np.rint(col) + rng.integers(-8, 9, size=(n, 3)), 0, 255).astype(np.uint8)
    perm = rng.permutation(n)
    return pack_points(xyz[perm], col[perm])


def gen_uniform(n, seed=0):
    """G-unif(N, seed): xyz uniform in [1/7, 6/7)^3, rgb uniform (worst case)."""
    rng = np.random.default_rng(seed)
    xyz = (1.0 / 7.0 + rng.random((n, 3)) * (5.0 / 7.0)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    return pack_points(xyz, rgb)


def gen_gof(n, seed=0, frames=30):
    """GOF(seed, F): a group of frames of the G-surf body under a slow rigid motion (BASELINE configs[2]):
    frame f is the frame-0 surface rotated by f * 0.4 degrees about the vertical axis through the body and shifted by
    f * 0.002 along x (a quarter of a 16-voxel macroblock per frame at 11 bits), with 1 % of its points re-sampled and every
    point's colour noise re-drawn (seed + f), in a fresh shuffled order.  All frames share ONE normalisation (the reference
    normalises a group with a common box, impl.hpp:1871-1986), so the motion survives it."""
    rng0 = np.random.default_rng(seed)
    a, b, c = _ELLIPSOIDS[:, 3], _ELLIPSOIDS[:, 4], _ELLIPSOIDS[:, 5]
    wts = a * b + a * c + b * c
    counts = rng0.multinomial(n, wts / wts.sum())
    parts = []
    for e, m in zip(_ELLIPSOIDS, counts):
        d = rng0.normal(size=(m, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        parts.append(e[:3] + d * e[3:])
    body = np.concatenate(parts)
    x, y, z = body[:, 0], body[:, 1], body[:, 2]
    base_col = np.stack([128 + 100 * np.sin(12 * x + 3 * z), 128 + 90 * np.cos(9 * y + 5 * z), 128 + 110 * np.sin(7 * z)], axis=1)
    clouds = []
    for f in range(frames):
        rng = np.random.default_rng(seed + 1000003 * (f + 1))
        th = np.deg2rad(0.4 * f)
        R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
        pts = body.copy()
        col = base_col.copy()
        k = max(1, n // 100)
        sel = rng.choice(n, k, replace=False)
        src = rng.choice(n, k)
        pts[sel] = body[src] + rng.normal(scale=1e-3, size=(k, 3))
        col[sel] = base_col[src]
        pts = (pts - [0.5, 0.5, 0.5]) @ R.T + [0.5, 0.5, 0.5] + [0.002 * f, 0, 0]
        colf = np.clip(np.rint(col) + rng.integers(-4, 5, size=(n, 3)), 0, 255).astype(np.uint8)
        perm = rng.permutation(n)
        clouds.append((pts[perm].astype(np.float32), colf[perm]))
    mn = np.min([c[0].min(axis=0) for c in clouds], axis=0)
    mx = np.max([c[0].max(axis=0) for c in clouds], axis=0)
    ext = (mx - mn).astype(np.float32)
    mn = (mn - np.float32(0.2) * ext).astype(np.float32)
    rngv = ((mx + np.float32(0.2) * ext).astype(np.float32) - mn).astype(np.float32)
    return [pack_points(((xyz - mn) / rngv).astype(np.float32), col) for xyz, col in clouds]
