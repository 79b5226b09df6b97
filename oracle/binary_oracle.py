"""CPU restatement of the reference's code bit-packing.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and by nothing in the product path).  Pinned against the
UNMODIFIED reference: ``oracle/gen_golden_bits.py`` runs ``/root/reference/binary.py`` (BitPacker / BitUnpacker)
through the loop of ``compress.compress_to_file`` and stores the byte streams under ``tests/golden/bitpack_*.npz``;
``tests/test_bitpack_oracle.py`` checks this file against them byte for byte.

Reference behaviour restated here:
  * binary.py:69-78  ``BitPacker.push``: the stream is little-endian in bits: value number i occupies stream bits
    [i*bits, (i+1)*bits), byte j holds stream bits [8j, 8j+8);
  * binary.py:80-87  ``BitPacker.flush``: a trailing partial byte is written with its unused high bits zero;
  * compress.py:70-92 the order of the pushes for one segment ``frame [1, K, T]``: for t in range(T): for k in range(K):
    push(frame[0, k, t]) -- time-major, codebook-minor;
  * binary.py:104-121 ``BitUnpacker.pull``: the inverse; trailing bits that do not fill a value are ignored (the reference
    may return up to 8 // bits "ghost" values from the padding of the last byte, binary.py:147-148; callers read K*T values).
"""
from __future__ import annotations

import numpy as np


def packed_nbytes(n_values: int, bits: int) -> int:
    return (n_values * bits + 7) // 8


def pack_values_loop(values, bits: int) -> bytes:
    """binary.py:69-87 literally (pure-Python big-int accumulator); small cases only."""
    cur, nb, out = 0, 0, bytearray()
    for v in values:
        cur += int(v) << nb
        nb += bits
        while nb >= 8:
            out.append(cur & 0xFF)
            nb -= 8
            cur >>= 8
    if nb:
        out.append(cur)
    return bytes(out)


def unpack_values_loop(data: bytes, bits: int, n_values: int):
    """binary.py:104-121 for the first ``n_values`` values."""
    cur, nb, pos, out = 0, 0, 0, []
    mask = (1 << bits) - 1
    for _ in range(n_values):
        while nb < bits:
            cur += data[pos] << nb
            pos += 1
            nb += 8
        out.append(cur & mask)
        cur >>= bits
        nb -= bits
    return out


def pack_values(values: np.ndarray, bits: int) -> np.ndarray:
    """Vectorised equivalent of ``pack_values_loop`` (numpy bit matrix), any size."""
    v = np.asarray(values, dtype=np.uint64).reshape(-1)
    n = v.size
    bit = ((v[:, None] >> np.arange(bits, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
    pad = (-bit.size) % 8
    if pad:
        bit = np.concatenate([bit, np.zeros(pad, dtype=np.uint8)])
    return np.packbits(bit.reshape(-1, 8), axis=1, bitorder="little").reshape(-1)[: packed_nbytes(n, bits)]


def unpack_values(data: np.ndarray, bits: int, n_values: int) -> np.ndarray:
    bit = np.unpackbits(np.asarray(data, dtype=np.uint8), bitorder="little")[: n_values * bits].reshape(n_values, bits)
    return (bit.astype(np.uint64) << np.arange(bits, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.int64)


def pack_frame(frame_bkt: np.ndarray, bits: int) -> np.ndarray:
    """One byte stream per batch item of ``frame [B, K, T]`` in the push order of compress.py:70-92 (t-major, k-minor).
    Returns uint8 ``[B, packed_nbytes(K*T, bits)]``."""
    f = np.asarray(frame_bkt)
    b, k, t = f.shape
    return np.stack([pack_values(f[i].T.reshape(-1), bits) for i in range(b)]) if b else np.zeros((0, packed_nbytes(k * t, bits)), np.uint8)


def unpack_frame(data_b: np.ndarray, k: int, t: int, bits: int) -> np.ndarray:
    d = np.asarray(data_b, dtype=np.uint8)
    return np.stack([unpack_values(d[i], bits, k * t).reshape(t, k).T for i in range(d.shape[0])]) if d.shape[0] else np.zeros((0, k, t), np.int64)
