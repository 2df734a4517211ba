"""Synthetic unit-norm embeddings, reproducible bit-for-bit on CPU and GPU.  TEST INFRASTRUCTURE.

SURVEY.md section 8d asks for a counter-based generator so that row ``r`` of the
gallery depends only on (seed, r): the same 1 M / 100 M gallery can then be
materialised shard by shard on any number of GPUs and re-derived on the host for
spot checks.  The device twin is ``frg_store_fill_synthetic`` (csrc/synth.cu).

Generator ``frg-synth-v1``
  * Philox4x32-10, key = (seed lo, seed hi), counter = (row lo, row hi, block, stream);
    one call yields 128 bits = 8 elements.
  * element = (sum of the four 4-bit nibbles of its 16-bit lane) - 30: a 4-term
    Irwin-Hall integer in [-30, 30], symmetric, sigma = sqrt(85).  Integer-valued on
    purpose: sum(x*x) over a row is an exact fp32 integer (< 2**24 for D <= 18641) in
    ANY summation order, and sqrt / divide are correctly rounded in both numpy and
    CUDA, so the normalised row is bit-identical on both sides without transcendental
    functions.
  * row = x / sqrt(sum(x*x)) in fp32.

Streams: 0 = gallery rows, 1 = impostor queries, 2 = noise added to genuine queries.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_SHIFT32 = np.uint64(32)

STREAM_GALLERY = 0
STREAM_IMPOSTOR = 1
STREAM_NOISE = 2

SIGMA = float(np.sqrt(85.0))  # std of one raw integer element

GALLERY_SEED = 1234
QUERY_SEED = 4321


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10.  c* are uint32 arrays (broadcastable); k* python ints."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0  # 32x32 -> 64, no overflow in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> _SHIFT32, p0 & _MASK32
        hi1, lo1 = p1 >> _SHIFT32, p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def raw_rows(rows, dim: int, seed: int, stream: int) -> np.ndarray:
    """Integer-valued fp32 rows (before normalisation) for the given global row indices."""
    if dim % 8:
        raise ValueError("dim must be a multiple of 8")
    rows = np.asarray(rows, dtype=np.int64).reshape(-1)
    nblk = dim // 8
    r_lo = (rows & 0xFFFFFFFF).astype(np.uint32)[:, None]
    r_hi = ((rows >> 32) & 0xFFFFFFFF).astype(np.uint32)[:, None]
    blk = np.arange(nblk, dtype=np.uint32)[None, :]
    w = philox4x32_10(r_lo, r_hi, blk, np.uint32(stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=-1)  # [n, nblk, 4] uint32
    lanes = np.stack([words & 0xFFFF, words >> 16], axis=-1)  # [n, nblk, 4, 2] (e = 2*word + half)
    lanes = lanes.reshape(len(rows), nblk * 8).astype(np.int32)
    s = (lanes & 0xF) + ((lanes >> 4) & 0xF) + ((lanes >> 8) & 0xF) + ((lanes >> 12) & 0xF)
    return (s - 30).astype(np.float32)


def unit_rows(rows, dim: int, seed: int = GALLERY_SEED, stream: int = STREAM_GALLERY) -> np.ndarray:
    """Unit-norm fp32 rows; bit-identical to the device generator."""
    x = raw_rows(rows, dim, seed, stream)
    ss = np.sum(x.astype(np.float64) ** 2, axis=1)  # exact integers
    norm = np.sqrt(ss.astype(np.float32))  # correctly rounded fp32 sqrt of an exact fp32 integer
    with np.errstate(divide="ignore", invalid="ignore"):
        return (x / norm[:, None]).astype(np.float32)


def gallery(n: int, dim: int, seed: int = GALLERY_SEED, row0: int = 0, chunk: int = 65536) -> np.ndarray:
    out = np.empty((n, dim), dtype=np.float32)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        out[a:b] = unit_rows(np.arange(row0 + a, row0 + b), dim, seed, STREAM_GALLERY)
    return out


def queries(f: int, n_gallery: int, dim: int, seed: int = QUERY_SEED, gallery_seed: int = GALLERY_SEED,
            noise: float = 0.03, genuine_every: int = 2, q0: int = 0):
    """SURVEY.md section 8d query mix: even query index = genuine (gallery row + noise, NOT
    re-normalised here - the matcher normalises, as the reference does), odd = impostor.

    Returns (Q float32[f, dim], target int64[f]) with target = -1 for impostors.
    """
    qi = np.arange(q0, q0 + f, dtype=np.int64)
    impostor = unit_rows(qi, dim, seed, STREAM_IMPOSTOR)
    target = np.full(f, -1, dtype=np.int64)
    if n_gallery <= 0:
        return impostor, target
    # target row: a Philox draw keyed on the query index (stream 1, block index past the row data)
    w = philox4x32_10((qi & 0xFFFFFFFF).astype(np.uint32), np.uint32(0), np.uint32(0xFFFFFFFF),
                      np.uint32(STREAM_IMPOSTOR), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    pick = ((w[0].astype(np.uint64) << np.uint64(32)) | w[1].astype(np.uint64)) % np.uint64(n_gallery)
    genuine = (qi % genuine_every) == 0
    target[genuine] = pick[genuine].astype(np.int64)
    out = impostor.copy()
    if genuine.any():
        g = unit_rows(target[genuine], dim, gallery_seed, STREAM_GALLERY)
        nz = raw_rows(qi[genuine], dim, seed, STREAM_NOISE) / np.float32(SIGMA)
        # noise * N(0,1) per component at D=512 (SURVEY: cos ~ 1/sqrt(1 + noise^2 * 512) ~ 0.83);
        # scaled by sqrt(512/D) so other dims keep the same angle
        per_component = np.float32(noise * np.sqrt(512.0 / dim))
        out[genuine] = (g + per_component * nz).astype(np.float32)
    return out.astype(np.float32), target
