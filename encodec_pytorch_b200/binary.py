"""Device-side code bit-packing: the byte streams of the reference's ``binary.BitPacker`` / ``BitUnpacker``
(binary.py:55-121) for whole frames at once.

The reference packs a segment's codes with a Python loop, one ``push`` per value and one ``fo.write`` per byte
(compress.py:70-92: ``for t in range(T): for k, value in enumerate(frame[0, :, t].tolist()): packer.push(value)``).
Here the same stream (time-major, codebook-minor, little-endian in bits, zero-padded last byte) is produced for every
batch item by one kernel launch (``rvq_bitpack``), so ``compress_to_file`` becomes
``fo.write(pack_frame(frame, bits)[0].cpu().numpy().tobytes())``; ``unpack_frame`` is the inverse for
``decompress_from_file`` (compress.py:128-147).  CUDA tensors only, like the rest of the package.
"""
from __future__ import annotations

import torch

from . import _lib as L


def packed_nbytes(n_codebooks: int, n_steps: int, bits: int) -> int:
    """Bytes of one stream: ``ceil(K * T * bits / 8)`` (BitPacker.flush pads the last byte, binary.py:80-87)."""
    return (n_codebooks * n_steps * bits + 7) // 8


def _check(bits: int) -> None:
    if not 1 <= int(bits) <= 16:
        raise RuntimeError(f"bits per codebook must be in 1..16, got {bits}")


def pack_frame(frame: torch.Tensor, bits: int) -> torch.Tensor:
    """``frame``: int64 CUDA tensor ``[B, K, T]`` (any strides; ``model.encode``'s frames are transposed views of the
    search's ``[K, B, T]`` output).  Returns uint8 ``[B, packed_nbytes(K, T, bits)]``: row b is the byte stream
    ``BitPacker(bits, fo)`` writes for ``frame[b]`` followed by ``flush()``."""
    _check(bits)
    if not frame.is_cuda or frame.dtype != torch.int64 or frame.dim() != 3:
        raise RuntimeError("pack_frame: expected a CUDA int64 tensor [B, K, T]; there is no CPU path")
    B, K, T = (int(v) for v in frame.shape)
    nbytes = packed_nbytes(K, T, bits)
    out = torch.empty((B, nbytes), dtype=torch.uint8, device=frame.device)
    if B == 0 or nbytes == 0:
        return out
    sb, sk, st = (int(v) for v in frame.stride())
    lib = L.load()
    with torch.cuda.device(frame.device):
        L.check(lib.rvq_bitpack(frame.data_ptr(), sk, sb, st, K, B, T, int(bits), out.data_ptr(), nbytes,
                                L.stream_ptr(frame.device)), "rvq_bitpack")
    return out


def unpack_frame(data: torch.Tensor, n_codebooks: int, n_steps: int, bits: int) -> torch.Tensor:
    """Inverse of :func:`pack_frame`: ``data`` uint8 CUDA ``[B, >= packed_nbytes]`` -> int64 ``[B, K, T]``
    (the values ``BitUnpacker(bits, fo).pull()`` returns, in the order compress.py:139-146 consumes them)."""
    _check(bits)
    if not data.is_cuda or data.dtype != torch.uint8 or data.dim() != 2:
        raise RuntimeError("unpack_frame: expected a CUDA uint8 tensor [B, nbytes]; there is no CPU path")
    B = int(data.shape[0])
    nbytes = packed_nbytes(n_codebooks, n_steps, bits)
    if int(data.shape[1]) < nbytes:
        raise RuntimeError(f"unpack_frame: {data.shape[1]} bytes per stream, {nbytes} needed")
    data = data.contiguous()
    frame = torch.empty((B, n_codebooks, n_steps), dtype=torch.int64, device=data.device)
    if B == 0 or nbytes == 0:
        return frame
    sb, sk, st = (int(v) for v in frame.stride())
    lib = L.load()
    with torch.cuda.device(data.device):
        L.check(lib.rvq_bitunpack(data.data_ptr(), int(data.stride(0)), n_codebooks, B, n_steps, int(bits), frame.data_ptr(),
                                  sk, sb, st, L.stream_ptr(data.device)), "rvq_bitunpack")
    return frame
