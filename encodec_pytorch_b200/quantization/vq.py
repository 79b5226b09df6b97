"""Host-side mirror of the reference's ``quantization/vq.py``: ``QuantizedResult`` (vq.py:19-25)
and ``ResidualVectorQuantizer`` (vq.py:28-128) with identical signatures, attributes and
``state_dict`` keys (``vq.layers.{i}._codebook.{inited,cluster_size,embed,embed_avg}``), so that
``model.py``'s ``EncodecModel``, ``compress.py`` and ``train_multi_gpu.py`` can use it unchanged.
The arithmetic runs in the sm_100a kernels behind ``include/rvq_b200.h``."""
from __future__ import annotations

from dataclasses import dataclass, field
import math
import typing as tp

import torch
from torch import nn

from .core_vq import ResidualVectorQuantization


@dataclass
class QuantizedResult:
    """vq.py:19-25."""
    quantized: torch.Tensor
    codes: torch.Tensor
    bandwidth: torch.Tensor  # bandwidth in kb/s used, per batch item.
    penalty: tp.Optional[torch.Tensor] = None
    metrics: dict = field(default_factory=dict)


class ResidualVectorQuantizer(nn.Module):
    """Residual Vector Quantizer (vq.py:28-128).

    Args (same defaults as the reference, vq.py:56-65):
        dimension, n_q, bins, decay, kmeans_init, kmeans_iters, threshold_ema_dead_code.
    """

    def __init__(self, dimension: int = 256, n_q: int = 8, bins: int = 1024, decay: float = 0.99,
                 kmeans_init: bool = True, kmeans_iters: int = 50, threshold_ema_dead_code: int = 2):
        super().__init__()
        self.n_q = n_q
        self.dimension = dimension
        self.bins = bins
        self.decay = decay
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.vq = ResidualVectorQuantization(
            dim=self.dimension,
            codebook_size=self.bins,
            num_quantizers=self.n_q,
            decay=self.decay,
            kmeans_init=self.kmeans_init,
            kmeans_iters=self.kmeans_iters,
            threshold_ema_dead_code=self.threshold_ema_dead_code,
        )

    def forward(self, x: torch.Tensor, sample_rate: int, bandwidth: tp.Optional[float] = None) -> QuantizedResult:
        """vq.py:84-99.  ``sample_rate`` is what the callers pass there: the FRAME rate
        (model.py:207)."""
        bw_per_q = self.get_bandwidth_per_quantizer(sample_rate)
        n_q = self.get_num_quantizers_for_bandwidth(sample_rate, bandwidth)
        quantized, codes, commit_loss = self.vq(x, n_q=n_q)
        # == torch.tensor(n_q * bw_per_q).to(x) (vq.py:94), built on the device: the pageable host-to-device copy of the
        # reference form synchronises the stream
        bw = torch.full((), n_q * bw_per_q, dtype=x.dtype, device=x.device)
        return QuantizedResult(quantized, codes, bw, penalty=torch.mean(commit_loss))

    def get_num_quantizers_for_bandwidth(self, sample_rate: int, bandwidth: tp.Optional[float] = None) -> int:
        """vq.py:101-108: all stages when ``bandwidth`` is falsy, else ``max(1, floor(bw / bw_per_q))``
        (not capped here; the stack caps at its length)."""
        bw_per_q = self.get_bandwidth_per_quantizer(sample_rate)
        n_q = self.n_q
        if bandwidth and bandwidth > 0.:
            n_q = int(max(1, math.floor(bandwidth / bw_per_q)))
        return n_q

    def get_bandwidth_per_quantizer(self, sample_rate: int):
        """vq.py:110-113."""
        return math.log2(self.bins) * sample_rate / 1000

    def encode(self, x: torch.Tensor, sample_rate: int, bandwidth: tp.Optional[float] = None,
               layout: str = "kbt") -> torch.Tensor:
        """vq.py:115-122: ``[B, D, T]`` fp32 -> ``[n_q, B, T]`` int64 (``layout="bkt"``: the contiguous ``[B, n_q, T]``
        tensor model.py:166 builds with a transpose)."""
        n_q = self.get_num_quantizers_for_bandwidth(sample_rate, bandwidth)
        return self.vq.encode(x, n_q=n_q, layout=layout)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """vq.py:124-128: ``[n_q, B, T]`` int64 -> ``[B, D, T]`` fp32."""
        return self.vq.decode(codes)

    # ---- the 48 kHz model's segment loop (model.py:141-145 / :178-179) in one launch --------------------------------
    def encode_segments(self, segments: tp.Sequence[torch.Tensor], sample_rate: int,
                        bandwidth: tp.Optional[float] = None, layout: str = "kbt") -> tp.List[torch.Tensor]:
        """``[self.encode(s, sample_rate, bandwidth) for s in segments]`` for latents ``[B, D, T_i]`` of one batch size,
        with ONE fused launch over all segments: frames are independent (core_vq.py works row-wise on ``[N, D]``), so
        the segments are laid end to end along time, searched together and the codes handed back as per-segment views
        (``[n_q, B, T_i]``; ``[B, n_q, T_i]`` with ``layout="bkt"``).  The reference encodes the 31 one-second segments
        of a 30 s clip one call at a time (4 800 frames each: a quarter of a B200 per launch)."""
        if len(segments) == 0:
            return []
        if len(segments) == 1:
            return [self.encode(segments[0], sample_rate, bandwidth, layout)]
        b = segments[0].shape[0]
        if any(s.dim() != 3 or s.shape[0] != b or s.shape[1] != self.dimension for s in segments):
            raise RuntimeError("encode_segments: expected latents [B, D, T_i] with a common batch size")
        lens = [int(s.shape[2]) for s in segments]
        codes = self.encode(torch.cat(list(segments), dim=2), sample_rate, bandwidth, layout)
        return list(torch.split(codes, lens, dim=2))

    def decode_segments(self, codes: tp.Sequence[torch.Tensor]) -> tp.List[torch.Tensor]:
        """``[self.decode(c) for c in codes]`` (codes ``[n_q, B, T_i]``, any strides) with one launch."""
        if len(codes) <= 1:
            return [self.decode(c) for c in codes]
        lens = [int(c.shape[2]) for c in codes]
        out = self.decode(torch.cat(list(codes), dim=2))
        return list(torch.split(out, lens, dim=2))

    @property
    def contiguous_outputs(self) -> bool:
        """See ``ResidualVectorQuantization.contiguous_outputs``."""
        return self.vq.contiguous_outputs

    @contiguous_outputs.setter
    def contiguous_outputs(self, value: bool) -> None:
        self.vq.contiguous_outputs = bool(value)

    def invalidate(self) -> None:
        """Forget the cached search image; call after writing codebook buffers through ``.data``."""
        self.vq.invalidate()
