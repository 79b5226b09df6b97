"""Seeded synthetic inputs shared by the golden-vector generator and the tests.

TEST INFRASTRUCTURE ONLY (see oracle/rvq_oracle.py header).  All inputs are
regenerated from seeds with the torch CPU generator (bit-stable within this
image's torch build); the fixtures keep a sha256 of every regenerated input so
a drifting RNG is detected instead of silently comparing different tensors.
"""
from __future__ import annotations

import hashlib
import typing as tp

import torch

from . import rvq_oracle as O


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def latents(b: int, d: int, t: int, seed: int = 1234, scale: float = 1.0) -> torch.Tensor:
    """unit-variance fp32 latents ``[B, D, T]`` from a private generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, d, t, generator=g, dtype=torch.float32) * scale


def codebooks(d: int, k: int, n_q: int, seed: int = 0, kmeans_init: bool = False) -> tp.List[O.State]:
    """Same tables the reference constructor draws under ``torch.manual_seed(seed)``
    (kaiming-uniform per stage, stage 0 first)."""
    torch.manual_seed(seed)
    return O.new_rvq_states(d, k, n_q, kmeans_init)


class Case(tp.NamedTuple):
    name: str
    b: int
    d: int
    t: int
    k: int
    n_q: int
    frame_rate: float
    bandwidth: tp.Optional[float]
    x_seed: int = 1234
    cb_seed: int = 0
    x_scale: float = 1.0


# encode/decode/eval-forward cases (reference run in eval mode, kmeans_init=False)
ENCODE_CASES = [
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    Case("cfg1_b4_t750_nq8", 4, 128, 750, 1024, 8, 75, 6.0),
    # ragged / tiny shapes
    Case("tiny_b1_t1_nq2", 1, 128, 1, 1024, 2, 75, 1.5),
    Case("ragged_b3_t37_nq32", 3, 128, 37, 1024, 32, 75, 24.0),
    # 48 kHz-style stack (150 Hz frame rate, 16 stages), last short segment of model.py:141-145
    Case("seg48k_b2_t5_nq16", 2, 128, 5, 1024, 16, 150, 24.0),
    # generic small codebook (exercises the non-128/1024 exact kernel)
    Case("small_d16_k64_nq3", 2, 16, 37, 64, 3, 75, None, 99, 7),
    # bandwidth above what the stack supports: silently capped at len(layers)
    Case("capped_bw48_nq8", 1, 128, 20, 1024, 8, 75, 48.0, 5, 3),
    # large-magnitude latents (fp16 operand range / margin logic)
    Case("scaled_x30_nq8", 2, 128, 64, 1024, 8, 75, 6.0, 77, 1, 30.0),
]

# training-forward cases (kmeans_init=False so the first step is RNG-free except expiry)
TRAIN_CASES = [
    Case("train_b2_t300_nq4", 2, 128, 300, 1024, 4, 75, 3.0, 21, 4),
    Case("train_small_d16_k64_nq3", 2, 16, 150, 64, 3, 75, None, 22, 5),
]

# k-means init cases (init centroids injected so the RNG is not part of the contract)
KMEANS_CASES = [
    Case("kmeans_d16_k64_nq2", 2, 16, 250, 64, 2, 75, None, 31, 6),
    Case("kmeans_d128_k1024_nq2", 2, 128, 700, 1024, 2, 75, None, 32, 6),
]
