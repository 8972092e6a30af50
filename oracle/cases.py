"""Seeded input families shared by the golden-vector generator, the parity tests and
bench.py's synthetic workload (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every generator is a pure function of its seed (numpy ``default_rng`` / PCG64), so the
golden files only need to store outputs plus an input checksum.
"""
from __future__ import annotations

import hashlib

import numpy as np

K, H, W = 17, 64, 48


def planted_peak_logits(batch: int, seed: int = 0, amp=(2.0, 8.0), sigma=(1.0, 3.0), noise=0.05):
    """SURVEY §8(d) config-3 synthetic logits: a*exp(-r^2/2s^2) + N(0, noise), centres
    uniform over the map *including* the border rows/columns."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(-0.5, W - 0.5, (batch, K, 1, 1))
    cy = rng.uniform(-0.5, H - 0.5, (batch, K, 1, 1))
    a = rng.uniform(*amp, (batch, K, 1, 1))
    s = rng.uniform(*sigma, (batch, K, 1, 1))
    yy, xx = np.mgrid[0:H, 0:W]
    r2 = (xx[None, None] - cx) ** 2 + (yy[None, None] - cy) ** 2
    z = a * np.exp(-r2 / (2 * s**2)) + rng.normal(0, noise, (batch, K, H, W))
    return z.astype(np.float32)


def planted_peak_pair(batch: int, seed: int = 0, flip_indices=(0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15),
                      jitter: float = 0.05):
    """A flip-TTA pair like a real model produces: the flipped pass sees the mirrored image, so its map
    for keypoint flip_idx[k] is (up to noise) the mirror image of the plain pass's map for keypoint k.
    Returns (logits, logits_flipped_pass)."""
    z = planted_peak_logits(batch, seed)
    rng = np.random.default_rng(seed + 7919)
    inv = np.argsort(np.asarray(flip_indices))
    zf = z[:, inv][..., ::-1] * rng.uniform(0.9, 1.1, (batch, K, 1, 1)) + rng.normal(0, jitter, z.shape)
    return z, np.ascontiguousarray(zf.astype(np.float32))


def noise_logits(batch: int, seed: int, std: float):
    """Flat / random-init regime: logits ~ N(0, std)."""
    rng = np.random.default_rng(seed)
    return rng.normal(0, std, (batch, K, H, W)).astype(np.float32)


def uniform_heatmaps(batch: int, seed: int):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 1, (batch, K, H, W)).astype(np.float32)


def special_heatmaps():
    """Hand-built (2, K, H, W) edge cases: empty / constant maps, single pixels at
    corners and borders, exact 2-pixel plateau ties, near-border peaks."""
    hm = np.zeros((2, K, H, W), np.float32)
    p = hm[0]
    # k0: all zero (argmax must be flat index 0); k1: constant map
    p[1] = 0.25
    # single pixels: corners, borders, first interior cell
    for k, (y, x) in zip(range(2, 11), [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (0, 20), (30, 0), (H - 1, 7), (31, W - 1), (1, 1)]):
        p[k, y, x] = 1.0
    # exact ties: horizontal pair, vertical pair, 2x2 block, far-apart equal pixels
    p[11, 20, 10] = p[11, 20, 11] = 0.5
    p[12, 40, 30] = p[12, 41, 30] = 0.5
    p[13, 10:12, 40:42] = 0.25
    p[14, 5, 5] = p[14, 50, 40] = 0.5
    # two unequal blobs
    p[15, 12, 12] = 0.6
    p[15, 13, 13] = 0.4
    # peak one pixel from the right/bottom border (still interior)
    p[16, H - 2, W - 2] = 1.0
    q = hm[1]
    rng = np.random.default_rng(7)
    for k in range(K):  # a handful of random sparse pixels, like trained-model output
        n = int(rng.integers(1, 6))
        ys, xs = rng.integers(0, H, n), rng.integers(0, W, n)
        v = rng.uniform(0.05, 1, n).astype(np.float32)
        q[k, ys, xs] = v / v.sum()
    return hm


def checksum(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
