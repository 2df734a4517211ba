"""An adversarial case for the bf16 filter of the tensor-core variants (test infrastructure).

bf16 keeps 8 significand bits: round-to-nearest moves an element by up to 2^-8 relative (at the bottom of a
binade), so for unit vectors |q^.g^ - q.g| can reach 2 * 2^-8 + 2^-16 = 7.8e-3 when every rounding error lines
up - about twice what a "2^-9 per side" estimate gives.  Here they do line up:

  row A   first half of the dimensions, every element just BELOW a bf16 midpoint at a binade bottom: rounds down
  row B   second half, every element just ABOVE the midpoint: rounds up
  query   along A + B, its own elements placed the same way (down where A lives, up where B lives), norm 1 so
          that the matcher's re-normalisation leaves them in place

In exact arithmetic A beats B by 5e-4 (5x the 1e-4 id tolerance); in bf16 arithmetic B beats A by 8.6e-3.
A filter that keeps only rows within 2 * 4e-3 of the best bf16 score drops A and reports the wrong identity.
"""
import numpy as np


def bf16_rn(x):
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def adversarial_pair(seed=0, nfill=6):
    dim, h = 512, 256
    rng = np.random.default_rng(seed)
    sg = rng.choice([-1.0, 1.0], size=dim).astype(np.float32)
    mag = np.float32(2.0 ** -4)
    d = np.float32(1 + 0.98 * 2.0 ** -8)
    u = np.float32(1 + 1.02 * 2.0 ** -8)
    A = np.zeros(dim, np.float32)
    B = np.zeros(dim, np.float32)
    A[:h] = sg[:h] * mag * d
    B[h:] = sg[h:] * mag * u
    B[h] = sg[h] * np.float32(0.75) * mag * u              # one element at 3/4: exact A a little ahead of B
    step = 2.0 ** -12                                      # bf16 spacing in [2^-5, 2^-4)
    b = np.floor(2.0 ** -4.5 / step) * step
    q = np.zeros(dim, np.float64)
    q[:h] = sg[:h] * (b + 0.49 * step)
    q[h:] = sg[h:] * (b + 0.51 * step)
    fill = np.r_[np.arange(h - nfill, h), np.arange(dim - nfill, dim)]
    q[fill] = 0
    rem = 1.0 - np.sum(q ** 2)
    assert rem > 0
    q[fill] = sg[fill] * np.sqrt(rem / len(fill))          # ||q|| = 1: the re-normalisation is the identity
    return A, B, q.astype(np.float32)
