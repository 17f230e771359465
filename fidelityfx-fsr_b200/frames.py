"""Synthetic RGBA frames of SURVEY.md §8(c)/(d) — numpy only, used by tests, smoke and bench.

uniform(w,h,seed):     the LCG frame  s = s*1664525 + 1013904223 ; v = (s>>8)*2^-24  over row-major RGBA
structured(w,h,seed):  the same frame box-blurred 3x3, plus hard vertical / diagonal step edges and
                       64x64 blocks of exactly 0.0 and exactly 1.0 (exercises the `zro` branch, the
                       0*inf NaN paths of RCAS and the de-ringing clamps)
"""
import numpy as np

_A, _C = 1664525, 1013904223


def uniform(w, h, seed=12345):
    n = w * h * 4
    a = np.cumprod(np.full(n, _A, dtype=np.uint32), dtype=np.uint32)            # a^(i+1) mod 2^32
    geo = np.empty(n, dtype=np.uint32)                                           # sum_{k<=i} a^k
    geo[0] = 1
    if n > 1:
        geo[1:] = np.cumsum(a[:-1], dtype=np.uint32) + np.uint32(1)
    s = a * np.uint32(seed & 0xFFFFFFFF) + np.uint32(_C) * geo
    return ((s >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).reshape(h, w, 4)


def structured(w, h, seed=12345):
    f = uniform(w, h, seed).astype(np.float64)
    p = np.pad(f, ((1, 1), (1, 1), (0, 0)), mode="edge")
    blur = sum(p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)) / 9.0
    y, x = np.mgrid[0:h, 0:w]
    out = blur.copy()
    out[(x % 97) < 3] = 1.0                              # hard vertical bars
    out[((x + y) % 61) < 2] = 0.03125                    # diagonal steps
    out[((x - 2 * y) % 113) < 2, 1] = 0.875              # shallow-angle edge in green only
    bx, by = (x // 64), (y // 64)
    out[((bx + by) % 5 == 0)] = 0.0                      # flat black 64x64 blocks
    out[((bx + 2 * by) % 7 == 3)] = 1.0                  # flat white 64x64 blocks
    return out.astype(np.float32)


def to_half(f32):
    """Quantise to fp16 (round-to-nearest-even); the fp32 oracle then reads these quantised values."""
    return f32.astype(np.float16)
