"""Caller-side mirror of the reference's ``compress.py`` for the code path that feeds / consumes the RVQ:
``compress_to_file`` (compress.py:30-92) and ``decompress_from_file`` (compress.py:95-157) with the per-timestep Python
loops (``for t in range(T): for k, value in enumerate(frame[0, :, t].tolist()): packer.push(value)`` and its
``unpacker.pull()`` mirror) replaced by one ``pack_frame`` / ``unpack_frame`` launch per segment.  The byte stream is
identical to the reference's (`.ecdc` header ``'!4sBI'`` + JSON, one ``'!f'`` scale per normalised segment, then the
10-bit little-endian code stream, binary.py:17-53, :55-121).

Only the plain bit-packed stream is covered: ``use_lm=True`` (the LM entropy coder, model.py:27-65 + quantization/ac.py)
is outside this package and raises.  ``model`` is any object with the reference ``EncodecModel`` interface
(``encode``, ``decode``, ``name``, ``bits_per_codebook``, ``normalize``, ``segment_length``, ``segment_stride``,
``sample_rate``) -- in practice the reference's own class with ``qt`` swapped."""
from __future__ import annotations

import io
import json
import struct
import typing as tp

import torch

from .binary import pack_frame, packed_nbytes, unpack_frame

_HEADER = struct.Struct("!4sBI")          # binary.py:19: magic, protocol version, JSON length
_MAGIC = b"ECDC"


def write_ecdc_header(fo: tp.IO[bytes], metadata: tp.Any) -> None:
    """binary.py:23-29."""
    meta = json.dumps(metadata).encode("utf-8")
    fo.write(_HEADER.pack(_MAGIC, 0, len(meta)))
    fo.write(meta)
    fo.flush()


def _read_exactly(fo: tp.IO[bytes], size: int) -> bytes:
    """binary.py:32-42."""
    chunks = []
    while size > 0:
        got = fo.read(size)
        if not got:
            raise EOFError(f"Impossible to read enough data from the stream, {size} bytes remaining.")
        chunks.append(got)
        size -= len(got)
    return b"".join(chunks)


def read_ecdc_header(fo: tp.IO[bytes]) -> tp.Any:
    """binary.py:45-53."""
    magic, version, meta_size = _HEADER.unpack(_read_exactly(fo, _HEADER.size))
    if magic != _MAGIC:
        raise ValueError("File is not in ECDC format.")
    if version != 0:
        raise ValueError("Version not supported.")
    return json.loads(_read_exactly(fo, meta_size).decode("utf-8"))


def _no_lm(use_lm: bool) -> None:
    if use_lm:
        raise RuntimeError("the LM entropy coder (use_lm=True) is not part of the B200 RVQ path; use the plain stream")


def compress_to_file(model, wav: torch.Tensor, fo: tp.IO[bytes], use_lm: bool = False) -> None:
    """compress.py:30-92 for ``use_lm=False``: same header, same scales, same code bytes."""
    _no_lm(use_lm)
    assert wav.dim() == 2, "Only single waveform can be encoded."
    with torch.no_grad():
        frames = model.encode(wav[None])
    metadata = {
        "m": model.name,                  # model name
        "al": wav.shape[-1],              # audio_length
        "nc": frames[0][0].shape[1],      # num_codebooks
        "lm": use_lm,                     # use lm?
        "fr": frames[0][0].shape[2],
    }
    write_ecdc_header(fo, metadata)
    for frame, scale in frames:
        if scale is not None:
            fo.write(struct.pack("!f", scale.cpu().item()))
        fo.write(pack_frame(frame, model.bits_per_codebook)[0].cpu().numpy().tobytes())


def decompress_from_file(model, fo: tp.IO[bytes], device="cuda") -> tp.Tuple[torch.Tensor, int]:
    """compress.py:95-157 for streams written without the LM: returns ``(wav [C, T], sample_rate)``."""
    metadata = read_ecdc_header(fo)
    audio_length, num_codebooks = metadata["al"], metadata["nc"]
    assert isinstance(audio_length, int) and isinstance(num_codebooks, int)
    _no_lm(bool(metadata["lm"]))
    frames = []
    segment_length = model.segment_length or audio_length
    segment_stride = model.segment_stride or audio_length
    frame_length = metadata["fr"]                                  # compress.py:123: every segment carries "fr" steps
    nbytes = packed_nbytes(num_codebooks, frame_length, model.bits_per_codebook)
    for _offset in range(0, audio_length, segment_stride):
        scale = None
        if model.normalize:
            scale_f, = struct.unpack("!f", _read_exactly(fo, struct.calcsize("!f")))
            scale = torch.tensor(scale_f, device=device).view(1)
        raw = torch.frombuffer(bytearray(_read_exactly(fo, nbytes)), dtype=torch.uint8).to(device)
        frames.append((unpack_frame(raw[None], num_codebooks, frame_length, model.bits_per_codebook), scale))
    with torch.no_grad():
        wav = model.decode(frames)
    return wav[0, :, :audio_length], model.sample_rate


def compress(model, wav: torch.Tensor, use_lm: bool = False) -> bytes:
    """compress.py:160-175."""
    fo = io.BytesIO()
    compress_to_file(model, wav, fo, use_lm=use_lm)
    return fo.getvalue()


def decompress(model, compressed: bytes, device="cuda") -> tp.Tuple[torch.Tensor, int]:
    """compress.py:178-187."""
    return decompress_from_file(model, io.BytesIO(compressed), device=device)
