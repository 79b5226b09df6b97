"""Tensor-level wrappers over the C ABI (``include/rvq_b200.h``).

Every function takes CUDA tensors, allocates outputs with torch on the input's
device, and launches on the caller's current CUDA stream.  There is no
fallback: a non-CUDA / non-fp32 input or a failing call raises.
"""
from __future__ import annotations

import typing as tp

import torch

from . import _lib as L


class CodebookPack:
    """Search image of ``n_q`` codebooks (see ``rvq_pack`` in the header): one
    uint8 device buffer plus the shape it was built for."""

    __slots__ = ("buf", "n_q", "K", "D")

    def __init__(self, buf: torch.Tensor, n_q: int, K: int, D: int):
        self.buf, self.n_q, self.K, self.D = buf, n_q, K, D

    @property
    def device(self) -> torch.device:
        return self.buf.device


def _guard(device: torch.device):
    return torch.cuda.device(device)


def pack(embeds: tp.Sequence[torch.Tensor]) -> CodebookPack:
    """Build the search image of the given ``[K, D]`` fp32 tables (stage order)."""
    lib = L.load()
    n_q = len(embeds)
    e0 = embeds[0]
    K, D = int(e0.shape[0]), int(e0.shape[1])
    keep = []
    for e in embeds:
        L.require_cuda_f32(e, "codebook")
        if e.device != e0.device or tuple(e.shape) != (K, D):
            raise RuntimeError("all codebooks of a stack must share device and shape")
        keep.append(e if e.is_contiguous() else e.contiguous())
    nbytes = int(lib.rvq_pack_bytes(n_q, K, D))
    buf = torch.empty(nbytes, dtype=torch.uint8, device=e0.device)
    with _guard(e0.device):
        L.check(lib.rvq_pack(L.ptr_array(keep), n_q, K, D, buf.data_ptr(), nbytes, L.stream_ptr(e0.device)), "rvq_pack")
    return CodebookPack(buf, n_q, K, D)


def _strides_bdt(x: torch.Tensor) -> tp.Tuple[int, int, int]:
    sb, sd, st = x.stride()
    return int(sb), int(sd), int(st)


def encode(pk: CodebookPack, x: torch.Tensor, stage0: int, n_q: int, *,
           want_quantized: bool = False, want_sqerr: bool = False, want_residual: bool = False,
           quantized_accum: tp.Optional[torch.Tensor] = None, flags: int = 0,
           codes_bkt: bool = False, out_bdt: bool = False, ema_stats_out: tp.Optional[torch.Tensor] = None):
    """Fused multi-stage search on ``x [B, D, T]`` (any strides).

    Returns ``(codes [n_q,B,T] int64, quantized [B,T,D] | None, sqerr [n_q] float64 | None,
    residual [B,T,D] | None)``.  ``codes_bkt``: the codes come as a contiguous ``[B, n_q, T]`` tensor (model.py:166's
    layout) instead; ``out_bdt``: ``quantized`` comes as a contiguous ``[B, D, T]`` tensor.  ``ema_stats_out``: a flat
    fp32 buffer of ``n_q*K + n_q*K*D`` elements (``ema_stats_buffer``) that receives the EMA statistics of
    core_vq.py:227-228 from the same launch (``rvq_encode_train``)."""
    lib = L.load()
    L.require_cuda_f32(x, "x")
    if x.dim() != 3 or x.shape[1] != pk.D:
        raise RuntimeError(f"expected x of shape [B, {pk.D}, T], got {tuple(x.shape)}")
    if stage0 < 0 or stage0 + n_q > pk.n_q:
        raise RuntimeError("stage range outside the codebook pack")
    B, D, T = (int(v) for v in x.shape)
    dev = x.device
    if codes_bkt:
        flags |= L.FLAG_CODES_BKT
    if out_bdt:
        flags |= L.FLAG_OUT_BDT
    codes = torch.empty((B, n_q, T) if codes_bkt else (n_q, B, T), dtype=torch.int64, device=dev)
    qshape = (B, D, T) if out_bdt else (B, T, D)
    quantized = None
    if quantized_accum is not None:
        quantized = quantized_accum
        flags |= L.FLAG_ACCUM_Q
        assert quantized.is_contiguous() and tuple(quantized.shape) == qshape
    elif want_quantized:
        quantized = torch.empty(qshape, dtype=torch.float32, device=dev)
    sqerr = torch.zeros(n_q, dtype=torch.float64, device=dev) if want_sqerr else None
    residual = torch.empty((B, T, D), dtype=torch.float32, device=dev) if want_residual else None
    sb, sd, st = _strides_bdt(x)
    with _guard(dev):
        if ema_stats_out is not None:
            _, counts, esum = ema_stats_views(ema_stats_out, n_q, pk.K, pk.D)
            L.check(lib.rvq_encode_train(pk.buf.data_ptr(), pk.K, pk.D, x.data_ptr(), sb, sd, st, B, T, stage0, n_q,
                                         codes.data_ptr(), L.ptr(quantized), L.ptr(residual), L.ptr(sqerr),
                                         counts.data_ptr(), esum.data_ptr(), flags, L.stream_ptr(dev)), "rvq_encode_train")
        else:
            L.check(lib.rvq_encode(pk.buf.data_ptr(), pk.K, pk.D, x.data_ptr(), sb, sd, st, B, T, stage0, n_q,
                                   codes.data_ptr(), L.ptr(quantized), L.ptr(residual), L.ptr(sqerr), flags,
                                   L.stream_ptr(dev)), "rvq_encode")
    return codes, quantized, sqerr, residual


def ema_stats_buffer(n_q: int, K: int, D: int, device) -> torch.Tensor:
    """The flat fp32 buffer ``[n_q*K] counts || [n_q*K*D] embed_sum`` of the EMA statistics (one all-reduce covers it)."""
    return torch.empty(n_q * K + n_q * K * D, dtype=torch.float32, device=device)


def ema_stats_views(flat: torch.Tensor, n_q: int, K: int, D: int):
    n_cnt, n_sum = n_q * K, n_q * K * D
    if not (flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and flat.numel() >= n_cnt + n_sum):
        raise RuntimeError("EMA statistics buffer: expected a contiguous CUDA fp32 tensor of n_q*K*(D+1) elements")
    return flat, flat[:n_cnt].view(n_q, K), flat[n_cnt:n_cnt + n_sum].view(n_q, K, D)


def decode(pk: CodebookPack, codes: torch.Tensor, out_bdt: bool = False) -> torch.Tensor:
    """``codes [n_q, B, T]`` int64 (any strides) -> ``[B, T, D]`` fp32 (stage-ordered sum); a contiguous ``[B, D, T]``
    tensor with ``out_bdt``."""
    lib = L.load()
    if not codes.is_cuda:
        raise RuntimeError("decode: expected CUDA codes; the B200 RVQ path has no CPU fallback")
    if codes.dtype != torch.int64:
        codes = codes.to(torch.int64)
    if codes.dim() != 3:
        raise RuntimeError(f"expected codes of shape [n_q, B, T], got {tuple(codes.shape)}")
    n_q, B, T = (int(v) for v in codes.shape)
    if n_q > pk.n_q:
        raise RuntimeError(f"codes carry {n_q} stages, the stack has {pk.n_q}")
    out = torch.empty((B, pk.D, T) if out_bdt else (B, T, pk.D), dtype=torch.float32, device=codes.device)
    sq, sb, st = (int(v) for v in codes.stride())
    with _guard(codes.device):
        L.check(lib.rvq_decode_ex(pk.buf.data_ptr(), pk.K, pk.D, codes.data_ptr(), sq, sb, st, n_q, B, T,
                                  out.data_ptr(), L.FLAG_OUT_BDT if out_bdt else 0, L.stream_ptr(codes.device)), "rvq_decode_ex")
    return out


def ema_stats(pk: CodebookPack, x: torch.Tensor, codes: torch.Tensor, stage0: int, flags: int = 0,
              out: tp.Optional[torch.Tensor] = None) -> tp.Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-stage bincount and per-code sum of input residuals.  Returns ``(flat, counts [n_q,K],
    embed_sum [n_q,K,D])`` where both are views of the single fp32 buffer ``flat`` so that one
    all-reduce covers them."""
    lib = L.load()
    L.require_cuda_f32(x, "x")
    n_q, B, T = (int(v) for v in codes.shape)
    assert codes.is_contiguous() and codes.dtype == torch.int64
    K, D = pk.K, pk.D
    flat, counts, esum = ema_stats_views(out if out is not None else ema_stats_buffer(n_q, K, D, x.device), n_q, K, D)
    sb, sd, st = _strides_bdt(x)
    with _guard(x.device):
        L.check(lib.rvq_ema_stats(pk.buf.data_ptr(), K, D, x.data_ptr(), sb, sd, st, B, T, stage0, n_q,
                                  codes.data_ptr(), counts.data_ptr(), esum.data_ptr(), flags,
                                  L.stream_ptr(x.device)), "rvq_ema_stats")
    return flat, counts, esum


def ema_apply(cluster_sizes: tp.Sequence[torch.Tensor], embed_avgs: tp.Sequence[torch.Tensor],
              embeds: tp.Sequence[torch.Tensor], counts: torch.Tensor, embed_sum: torch.Tensor,
              decay: float, epsilon: float) -> None:
    lib = L.load()
    n_q = len(embeds)
    K, D = int(embeds[0].shape[0]), int(embeds[0].shape[1])
    for t in (*cluster_sizes, *embed_avgs, *embeds):
        L.require_cuda_f32(t, "codebook buffer")
        if not t.is_contiguous():
            raise RuntimeError("codebook buffers must be contiguous")
    dev = embeds[0].device
    with _guard(dev):
        L.check(lib.rvq_ema_apply(L.ptr_array(cluster_sizes), L.ptr_array(embed_avgs), L.ptr_array(embeds),
                                  n_q, K, D, counts.data_ptr(), embed_sum.data_ptr(), float(decay),
                                  float(epsilon), L.stream_ptr(dev)), "rvq_ema_apply")


def expire_replace(embed: torch.Tensor, cluster_size: torch.Tensor, samples: torch.Tensor, threshold: float) -> None:
    lib = L.load()
    for t in (embed, cluster_size, samples):
        L.require_cuda_f32(t, "expire_replace")
    K, D = int(embed.shape[0]), int(embed.shape[1])
    samples = samples.contiguous()
    assert tuple(samples.shape) == (K, D)
    with _guard(embed.device):
        L.check(lib.rvq_expire_replace(embed.data_ptr(), cluster_size.data_ptr(), samples.data_ptr(), K, D,
                                       float(threshold), L.stream_ptr(embed.device)), "rvq_expire_replace")


def expire_codes(pk: CodebookPack, x: torch.Tensor, codes: torch.Tensor, stage0: int, stage: int,
                 sel: torch.Tensor, cluster_size: torch.Tensor, threshold: float, embed: torch.Tensor,
                 flags: int = 0) -> None:
    lib = L.load()
    L.require_cuda_f32(x, "x")
    B, D, T = (int(v) for v in x.shape)
    assert codes.is_contiguous() and sel.dtype == torch.int64 and sel.is_contiguous() and sel.numel() == pk.K
    sb, sd, st = _strides_bdt(x)
    with _guard(x.device):
        L.check(lib.rvq_expire_codes(pk.buf.data_ptr(), pk.K, pk.D, x.data_ptr(), sb, sd, st, B, T, stage0, stage,
                                     codes.data_ptr(), sel.data_ptr(), cluster_size.data_ptr(), float(threshold),
                                     embed.data_ptr(), flags, L.stream_ptr(x.device)), "rvq_expire_codes")


def expire_stack(pk: CodebookPack, x: torch.Tensor, codes: torch.Tensor, stage0: int,
                 cluster_sizes: tp.Sequence[torch.Tensor], embeds: tp.Sequence[torch.Tensor], threshold: float,
                 seed: int, offset: int, flags: int = 0) -> tp.Tuple[torch.Tensor, torch.Tensor]:
    """Dead-code expiry of ``len(embeds)`` stages in two launches and no host sync (``rvq_expire_stack``).
    Returns ``(sel [n, K] int64, fired [n] int32)``: the drawn frame numbers and the per-stage "any code
    below threshold" flags, both on the device."""
    lib = L.load()
    L.require_cuda_f32(x, "x")
    B, D, T = (int(v) for v in x.shape)
    n = len(embeds)
    for t in (*cluster_sizes, *embeds):
        L.require_cuda_f32(t, "codebook buffer")
        if not t.is_contiguous():
            raise RuntimeError("codebook buffers must be contiguous")
    assert codes.is_contiguous() and len(cluster_sizes) == n
    sel = torch.empty((n, pk.K), dtype=torch.int64, device=x.device)
    fired = torch.empty((n,), dtype=torch.int32, device=x.device)
    sb, sd, st = _strides_bdt(x)
    with _guard(x.device):
        L.check(lib.rvq_expire_stack(pk.buf.data_ptr(), pk.K, pk.D, x.data_ptr(), sb, sd, st, B, T, stage0, n,
                                     codes.data_ptr(), L.ptr_array(cluster_sizes), L.ptr_array(embeds), float(threshold),
                                     int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), sel.data_ptr(), fired.data_ptr(),
                                     flags, L.stream_ptr(x.device)), "rvq_expire_stack")
    return sel, fired


def kmeans_assign(pk: CodebookPack, samples: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    L.require_cuda_f32(samples, "samples")
    samples = samples.contiguous()
    N = int(samples.shape[0])
    buckets = torch.empty(N, dtype=torch.int64, device=samples.device)
    with _guard(samples.device):
        L.check(lib.rvq_kmeans_assign(pk.buf.data_ptr(), pk.K, pk.D, samples.data_ptr(), N, buckets.data_ptr(),
                                      L.stream_ptr(samples.device)), "rvq_kmeans_assign")
    return buckets


def kmeans_update(samples: torch.Tensor, buckets: torch.Tensor, means: torch.Tensor) -> torch.Tensor:
    """In-place centroid update of ``means [K, D]``; returns ``bins [K]`` int64."""
    lib = L.load()
    L.require_cuda_f32(samples, "samples")
    samples = samples.contiguous()
    N, D = int(samples.shape[0]), int(samples.shape[1])
    K = int(means.shape[0])
    bins = torch.empty(K, dtype=torch.int64, device=samples.device)
    sums = torch.empty((K, D), dtype=torch.float32, device=samples.device)
    with _guard(samples.device):
        L.check(lib.rvq_kmeans_update(samples.data_ptr(), N, D, buckets.data_ptr(), K, means.data_ptr(),
                                      bins.data_ptr(), sums.data_ptr(), L.stream_ptr(samples.device)),
                "rvq_kmeans_update")
    return bins


def residual_combine(pk: CodebookPack, x: torch.Tensor, codes: torch.Tensor, stage0: int, w: torch.Tensor,
                     flags: int = 0) -> torch.Tensor:
    """``out [B, T, D] = sum_i w[i] * r_{i+1}`` (residual after stage i), recomputed from x and codes."""
    lib = L.load()
    L.require_cuda_f32(x, "x")
    B, D, T = (int(v) for v in x.shape)
    n_q = int(codes.shape[0])
    assert codes.is_contiguous() and codes.dtype == torch.int64
    w = w.to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty((B, T, D), dtype=torch.float32, device=x.device)
    sb, sd, st = _strides_bdt(x)
    with _guard(x.device):
        L.check(lib.rvq_residual_combine(pk.buf.data_ptr(), pk.K, pk.D, x.data_ptr(), sb, sd, st, B, T, stage0,
                                         n_q, codes.data_ptr(), w.data_ptr(), out.data_ptr(), flags,
                                         L.stream_ptr(x.device)), "rvq_residual_combine")
    return out


class search_counters:
    """Context manager: counts what the tcgen05 search does inside the block (``rvq_search_counters``).

        with ops.search_counters(device) as c:
            ops.encode(...)
        c.read()  ->  {"searched", "certified", "rescored", "fullscan"}   (synchronises)

    Off by default: the library writes no counters unless a buffer is registered for the calling thread."""

    NAMES = ("searched", "certified", "rescored", "fullscan")

    def __init__(self, device):
        self.buf = torch.zeros(32, dtype=torch.int64, device=device)

    def __enter__(self):
        L.check(L.load().rvq_search_counters(self.buf.data_ptr()), "rvq_search_counters")
        return self

    def __exit__(self, *exc):
        L.check(L.load().rvq_search_counters(None), "rvq_search_counters")
        return False

    def read(self) -> tp.Dict[str, int]:
        vals = self.buf.cpu().tolist()
        out = {n: int(vals[i]) for i, n in enumerate(self.NAMES)}
        out["wide"], out["wide_candidates"] = int(vals[11]), int(vals[12])     # frames with more than 4 candidates, their total
        return out


class pack_bound_mode:
    """Context manager: the score-error bound ``rvq_pack`` prepares the tcgen05 search with (``rvq_pack_bound_mode``):
    0 = chosen per stage (default), 1 = per-code bound on every stage, 2 = per-stage bound on every stage.  Both bounds
    are rigorous; packs built inside the block keep their mode (cached packs need ``invalidate()`` to be rebuilt)."""

    def __init__(self, mode: int):
        self.mode = int(mode)

    def __enter__(self):
        self.prev = L.load().rvq_pack_bound_mode(self.mode)
        return self

    def __exit__(self, *exc):
        L.load().rvq_pack_bound_mode(self.prev)
        return False
