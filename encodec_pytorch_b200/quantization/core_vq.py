"""Host-side mirror of the reference's ``quantization/core_vq.py`` for the B200 RVQ path.

Same classes, constructor arguments, buffer names and ``state_dict`` keys as
the reference (``EuclideanCodebook`` core_vq.py:105-237, ``VectorQuantization``
:240-324, ``ResidualVectorQuantization`` :327-375, helpers :45-102), but every
arithmetic step runs in the hand-written sm_100a kernels behind the C ABI
(``include/rvq_b200.h``) -- a residual stack is ONE fused launch instead of the
reference's per-stage ATen loop.  CUDA fp32 tensors only: anything else raises
(no CPU fallback, no other backend).
"""
from __future__ import annotations

import typing as tp
import warnings

import torch
from torch import nn

from .. import _lib as L
from .. import _ops as ops
from .. import distrib


def default(val: tp.Any, d: tp.Any) -> tp.Any:
    """core_vq.py:45-46."""
    return val if val is not None else d


def ema_inplace(moving_avg: torch.Tensor, new: torch.Tensor, decay: float) -> None:
    """core_vq.py:49-56 (kept for API parity; the stack update runs in ``rvq_ema_apply``)."""
    moving_avg.data.mul_(decay).add_(new, alpha=(1 - decay))


def laplace_smoothing(x: torch.Tensor, n_categories: int, epsilon: float = 1e-5) -> torch.Tensor:
    """core_vq.py:59-60."""
    return (x + epsilon) / (x.sum() + n_categories * epsilon)


def uniform_init(*shape: int) -> torch.Tensor:
    """core_vq.py:63-66."""
    t = torch.empty(shape)
    nn.init.kaiming_uniform_(t)
    return t


def sample_indices(num_samples: int, num: int, device: torch.device) -> torch.Tensor:
    """Index draw of core_vq.py:69-77: ``randperm`` prefix, or ``randint`` with replacement when
    there are fewer samples than requested.  Consumes the device's global torch RNG like the
    reference does."""
    if num_samples >= num:
        return torch.randperm(num_samples, device=device)[:num]
    return torch.randint(0, num_samples, (num,), device=device)


def sample_vectors(samples: torch.Tensor, num: int) -> torch.Tensor:
    """core_vq.py:69-77."""
    return samples[sample_indices(samples.shape[0], num, samples.device)]


class _LloydGraph:
    """One Lloyd iteration (pack, assign, scatter / finalise) on fixed buffers, captured once into a CUDA graph and replayed.
    The buffers belong to the graph: a call copies its samples and starting means in and the result out, so one graph serves
    every k-means of the same shape (the stages of a stack, later calls).  Capturing a fresh graph per call cost 2-60 ms each
    (growing with the number of graphs the process had made) against 5 ms for the 49 replays of a stage."""

    def __init__(self, n: int, k: int, d: int, device: torch.device):
        self.samples = torch.empty((n, d), dtype=torch.float32, device=device)
        self.means = torch.empty((k, d), dtype=torch.float32, device=device)
        self.graph: tp.Optional[torch.cuda.CUDAGraph] = None
        self.bins: tp.Optional[torch.Tensor] = None

    def step(self) -> torch.Tensor:
        pk = ops.pack([self.means])
        buckets = ops.kmeans_assign(pk, self.samples)
        return ops.kmeans_update(self.samples, buckets, self.means)

    def run(self, num_iters: int) -> torch.Tensor:
        if num_iters <= 0:
            return torch.zeros(self.means.shape[0], dtype=torch.int64, device=self.means.device)
        if torch.cuda.is_current_stream_capturing():     # already inside somebody's capture: plain launches (same kernels)
            for _ in range(num_iters):
                bins = self.step()
            return bins
        if self.graph is None:
            # the first iteration runs eagerly (it also pays the one-time function-attribute calls), the second is captured
            bins = self.step()
            num_iters -= 1
            if num_iters == 0:
                return bins
            # (the debug counters of rvq_search_counters are a per-thread pointer the launches read: a captured launch would
            # keep it beyond the life of the caller's buffer, so the capture runs with the counters off)
            L.check(L.load().rvq_search_counters(None), "rvq_search_counters")
            graph = torch.cuda.CUDAGraph()
            dev = self.means.device
            cur, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):           # (the raw capture API: torch.cuda.graph() would add a gc.collect per call)
                graph.capture_begin()
                self.bins = self.step()
                graph.capture_end()
            cur.wait_stream(side)
            self.graph = graph
        for _ in range(num_iters):
            self.graph.replay()
        return self.bins.clone()


_lloyd_graphs: tp.Dict[tp.Tuple[int, int, int, int], _LloydGraph] = {}


def kmeans(samples: torch.Tensor, num_clusters: int, num_iters: int = 10,
           init_means: tp.Optional[torch.Tensor] = None) -> tp.Tuple[torch.Tensor, torch.Tensor]:
    """core_vq.py:80-102: Lloyd iterations on flat ``[N, D]`` samples.  Assignment (direct
    sum-of-squared-differences, lowest index on ties) and the bincount / scatter-add centroid
    update (empty clusters keep their mean) are the ``rvq_kmeans_*`` kernels.  ``init_means``
    lets a caller inject the starting centroids instead of drawing them."""
    L.require_cuda_f32(samples, "kmeans samples")
    means0 = sample_vectors(samples, num_clusters) if init_means is None else init_means
    n, d = (int(v) for v in samples.shape)
    dev = samples.device
    key = (n, int(num_clusters), d, dev.index if dev.index is not None else torch.cuda.current_device())
    lg = _lloyd_graphs.get(key)
    if lg is None:
        _lloyd_graphs.clear()                     # one shape at a time: the buffers of an old shape are released
        lg = _lloyd_graphs[key] = _LloydGraph(n, int(num_clusters), d, dev)
    lg.samples.copy_(samples)
    lg.means.copy_(means0)
    bins = lg.run(num_iters)
    return lg.means.clone(), bins


def _flat_as_bdt(flat: torch.Tensor) -> torch.Tensor:
    """``[N, D]`` contiguous frames as the ``[1, D, N]`` strided view the kernels take."""
    return flat.t().unsqueeze(0)


class EuclideanCodebook(nn.Module):
    """Codebook with Euclidean distance (core_vq.py:105-237); see the reference for the argument
    documentation.  Buffers: ``inited [1]``, ``cluster_size [K]``, ``embed [K, D]``,
    ``embed_avg [K, D]`` (all fp32)."""

    def __init__(self, dim: int, codebook_size: int, kmeans_init: int = False, kmeans_iters: int = 10,
                 decay: float = 0.99, epsilon: float = 1e-5, threshold_ema_dead_code: int = 2):
        super().__init__()
        self.decay = decay
        init_fn: tp.Union[tp.Callable[..., torch.Tensor], tp.Any] = uniform_init if not kmeans_init else torch.zeros
        embed = init_fn(codebook_size, dim)

        self.codebook_size = codebook_size
        self.kmeans_iters = kmeans_iters
        self.epsilon = epsilon
        self.threshold_ema_dead_code = threshold_ema_dead_code

        self.register_buffer("inited", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(codebook_size))
        self.register_buffer("embed", embed)
        self.register_buffer("embed_avg", embed.clone())

        # host-side caches (not state): inited flag without a device sync, search image of `embed`
        self._inited_host: tp.Optional[bool] = None
        self._gen = 0
        self._pack_cache: tp.Optional[tp.Tuple[tp.Any, ops.CodebookPack]] = None
        # test hook: starting centroids for the next k-means init (instead of the RNG draw)
        self._kmeans_init_means: tp.Optional[torch.Tensor] = None

    # ---- cache management ---------------------------------------------------------------------
    def invalidate(self) -> None:
        """Forget host-side caches; call after mutating ``embed`` / ``inited`` through ``.data``."""
        self._gen += 1
        self._pack_cache = None
        self._inited_host = None

    def _tables_changed(self) -> None:
        self._gen += 1
        self._pack_cache = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate()

    def _is_inited(self) -> bool:
        """``if self.inited`` of core_vq.py:148 with the device read cached on the host."""
        if self._inited_host is None:
            self._inited_host = bool(self.inited.item())
        return self._inited_host

    def _key(self):
        return (self.embed.data_ptr(), self.embed._version, self._gen)

    def _pack(self) -> ops.CodebookPack:
        key = self._key()
        if self._pack_cache is None or self._pack_cache[0] != key:
            self._pack_cache = (key, ops.pack([self.embed]))
        return self._pack_cache[1]

    # ---- reference API ------------------------------------------------------------------------
    @torch.jit.ignore
    def init_embed_(self, data: torch.Tensor) -> None:
        """core_vq.py:146-157: k-means on the first batch; with ``distrib.sync_buffers(True)`` also the step the
        reference left as a FIXME: broadcast the initialised buffers from rank 0 so all ranks start in sync."""
        if self._is_inited():
            return
        L.require_cuda_f32(data, "init_embed_ data")
        with torch.no_grad():
            embed, cluster_size = kmeans(data, self.codebook_size, self.kmeans_iters, self._kmeans_init_means)
            self._kmeans_init_means = None
            self.embed.data.copy_(embed)
            self.embed_avg.data.copy_(embed.clone())
            self.cluster_size.data.copy_(cluster_size)
            self.inited.data.copy_(torch.Tensor([True]))
            if distrib.sync_enabled():             # opt-in (distrib.sync_buffers): all ranks start from rank 0's tables
                distrib.broadcast_tensors(self.buffers(), check=False)
        self._inited_host = True
        self._tables_changed()

    def replace_(self, samples: torch.Tensor, mask: torch.Tensor) -> None:
        """core_vq.py:159-163."""
        picked = sample_vectors(samples, self.codebook_size)
        with torch.no_grad():
            self.embed.data.copy_(torch.where(mask[..., None], picked, self.embed))
        self._tables_changed()

    def expire_codes_(self, batch_samples: torch.Tensor) -> None:
        """core_vq.py:165-175: rows whose EMA cluster size fell below the threshold are replaced by
        random batch rows (``rvq_expire_replace``).  ``torch.any`` is a host sync, as upstream."""
        if self.threshold_ema_dead_code == 0:
            return
        expired = self.cluster_size < self.threshold_ema_dead_code
        if not torch.any(expired):
            return
        flat = batch_samples.reshape(-1, batch_samples.shape[-1])
        picked = sample_vectors(flat, self.codebook_size)
        ops.expire_replace(self.embed, self.cluster_size, picked, float(self.threshold_ema_dead_code))
        self._tables_changed()

    def preprocess(self, x: torch.Tensor) -> torch.Tensor:
        """core_vq.py:177-179."""
        return x.reshape(-1, x.shape[-1])

    def quantize(self, x: torch.Tensor) -> torch.Tensor:
        """core_vq.py:181-189 on flat ``[N, D]`` frames: index of the nearest code."""
        L.require_cuda_f32(x, "quantize input")
        flat = x.contiguous()
        codes, _, _, _ = ops.encode(self._pack(), _flat_as_bdt(flat), 0, 1)
        return codes.view(-1)

    def postprocess_emb(self, embed_ind: torch.Tensor, shape) -> torch.Tensor:
        """core_vq.py:191-192."""
        return embed_ind.view(*shape[:-1])

    def dequantize(self, embed_ind: torch.Tensor) -> torch.Tensor:
        """core_vq.py:194-196: table gather (``rvq_decode`` with one stage)."""
        shape = tuple(embed_ind.shape)
        out = ops.decode(self._pack(), embed_ind.reshape(1, 1, -1))
        return out.view(*shape, self.embed.shape[1])

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """core_vq.py:198-206."""
        shape = x.shape
        return self.postprocess_emb(self.quantize(self.preprocess(x)), shape)

    def decode(self, embed_ind: torch.Tensor) -> torch.Tensor:
        """core_vq.py:208-210."""
        return self.dequantize(embed_ind)

    def forward(self, x: torch.Tensor) -> tp.Tuple[torch.Tensor, torch.Tensor]:
        """core_vq.py:212-237: init if needed, search + gather with the PRE-update table, then (in
        training) expiry and the EMA update."""
        shape = x.shape
        L.require_cuda_f32(x, "codebook input")
        flat = self.preprocess(x.detach()).contiguous()
        self.init_embed_(flat)
        pk = self._pack()
        codes, quant, _, _ = ops.encode(pk, _flat_as_bdt(flat), 0, 1, want_quantized=True)
        embed_ind = self.postprocess_emb(codes.view(-1), shape)
        quantize = quant.view(*shape)
        if self.training:
            with torch.no_grad():
                self.expire_codes_(flat)
                _update_stack([self], pk, _flat_as_bdt(flat), codes, 0, flags=0)
        return quantize, embed_ind


def _update_stack(codebooks: tp.Sequence[EuclideanCodebook], pk: ops.CodebookPack, x_bdt: torch.Tensor,
                  codes: torch.Tensor, stage0: int, flags: int, stats: tp.Optional[torch.Tensor] = None) -> None:
    """EMA update of core_vq.py:227-235 for a run of stages at once: bincount + per-code residual
    sums (``stats``: already accumulated by the search, ``rvq_encode_train``; else ``rvq_ema_stats`` from the codes),
    with ``distrib.sync_buffers(True)`` ONE all-reduce of the packed statistics
    across the frame shards (the role of distrib.all_reduce, distrib.py:32-34), then EMA / Laplace smoothing /
    table overwrite in place (``rvq_ema_apply``)."""
    if stats is not None:
        flat, counts, esum = ops.ema_stats_views(stats, len(codebooks), pk.K, pk.D)
    else:
        flat, counts, esum = ops.ema_stats(pk, x_bdt, codes, stage0, flags)
    if distrib.sync_enabled():                     # opt-in (distrib.sync_buffers): statistics of the global batch
        distrib.all_reduce_stats(flat)
    cb0 = codebooks[0]
    ops.ema_apply([cb.cluster_size for cb in codebooks], [cb.embed_avg for cb in codebooks],
                  [cb.embed for cb in codebooks], counts, esum, cb0.decay, cb0.epsilon)
    for cb in codebooks:
        cb._tables_changed()


class _AttachGrad(torch.autograd.Function):
    """Wires the already computed training outputs into autograd with the reference's gradient
    (SURVEY.md 3.4-6): every stage's straight-through term passes identity and residuals subtract
    detached values, so ``d quantized / d x = n_q * I``; ``d loss_i / d x = 2 w (r_i - ste_i) / n``
    (core_vq.py:309, :319-320, :348-349).  The residuals are recomputed from ``x`` and the codes
    with the pre-update tables held by ``pk`` (``rvq_residual_combine``)."""

    @staticmethod
    def forward(ctx, x, quantized, losses, codes, pk, n_q, commitment_weight, flags):
        ctx.save_for_backward(x, codes)
        ctx.pk, ctx.n_q, ctx.cw, ctx.flags = pk, n_q, commitment_weight, flags
        return quantized.view_as(quantized), losses.view_as(losses)

    @staticmethod
    def backward(ctx, g_quantized, g_losses):
        x, codes = ctx.saved_tensors
        grad = None
        if g_quantized is not None:
            grad = g_quantized * float(ctx.n_q)
        if g_losses is not None and ctx.cw > 0:
            b, d, t = x.shape
            w = g_losses.reshape(-1).to(torch.float32) * (2.0 * ctx.cw / float(b * t * d))
            comb = ops.residual_combine(ctx.pk, x, codes, 0, w, ctx.flags).permute(0, 2, 1)
            grad = comb if grad is None else grad + comb
        return grad, None, None, None, None, None, None, None


class VectorQuantization(nn.Module):
    """Vector quantization (core_vq.py:240-324); only Euclidean distance, like the reference."""

    def __init__(self, dim: int, codebook_size: int, codebook_dim: tp.Optional[int] = None, decay: float = 0.99,
                 epsilon: float = 1e-5, kmeans_init: bool = True, kmeans_iters: int = 50,
                 threshold_ema_dead_code: int = 2, commitment_weight: float = 1.):
        super().__init__()
        _codebook_dim: int = default(codebook_dim, dim)
        requires_projection = _codebook_dim != dim
        self.project_in = (nn.Linear(dim, _codebook_dim) if requires_projection else nn.Identity())
        self.project_out = (nn.Linear(_codebook_dim, dim) if requires_projection else nn.Identity())
        self.epsilon = epsilon
        self.commitment_weight = commitment_weight
        self._codebook = EuclideanCodebook(dim=_codebook_dim, codebook_size=codebook_size,
                                           kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                                           decay=decay, epsilon=epsilon,
                                           threshold_ema_dead_code=threshold_ema_dead_code)
        self.codebook_size = codebook_size

    @property
    def codebook(self) -> torch.Tensor:
        return self._codebook.embed

    @property
    def _plain(self) -> bool:
        return isinstance(self.project_in, nn.Identity) and isinstance(self.project_out, nn.Identity)

    def _require_plain(self) -> None:
        """The EnCodec models never project (model.py:260-264 builds ``dimension == codebook_dim``); the fused
        kernels cover exactly that case and nothing routes around them."""
        if not self._plain:
            raise RuntimeError("projected codebooks (codebook_dim != dim) are not supported by the B200 RVQ path")

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """core_vq.py:289-293: ``[B, D, T]`` -> ``[B, T]`` int64."""
        self._require_plain()
        L.require_cuda_f32(x, "x")
        codes, _, _, _ = ops.encode(self._codebook._pack(), x, 0, 1)
        return codes[0]

    def decode(self, embed_ind: torch.Tensor) -> torch.Tensor:
        """core_vq.py:295-299: ``[B, T]`` -> ``[B, D, T]`` (a permuted ``[B, T, D]`` buffer)."""
        self._require_plain()
        return self._codebook.decode(embed_ind).permute(0, 2, 1)

    def forward(self, x: torch.Tensor):
        """core_vq.py:301-324: returns ``(quantize [B, D, T], embed_ind [B, T], loss [1])``."""
        self._require_plain()
        quantized, codes, losses = _stack_forward([self], x, 1, self.training)
        return quantized, codes[0], losses[0]


def _warn_issue25() -> None:
    warnings.warn('When using RVQ in training model, first check '
                  'https://github.com/facebookresearch/encodec/issues/25 . '
                  'The bug wasn\'t fixed here for reproducibility.')


# Run the (provably overwritten) dead-code expiry kernels inside the training forward anyway -- for fidelity experiments.
RUN_DEAD_EXPIRY = False


def _rng_stream(device: torch.device, n: int) -> tp.Tuple[int, int]:
    """(seed, offset) of the device's global torch generator for a device-side draw, advancing it like a
    torch random op would (``torch.manual_seed`` therefore reproduces the draws)."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, off = gen.initial_seed(), gen.get_offset()
    gen.set_offset(off + 4 * max(1, n))
    return seed, off


def _expire_stack(layers: tp.Sequence[VectorQuantization], pk: ops.CodebookPack, x: torch.Tensor,
                  codes: torch.Tensor, stage0: int, flags: int) -> None:
    """Dead-code expiry (core_vq.py:165-175) for a run of stages.  The reference reads ``torch.any(expired)``
    on the host and draws ``randperm`` once per stage; here the "any" test, the index draw (same law:
    K distinct uniformly random frames in random order) and the row replacement of every stage run on the
    device in two launches, with no host sync (``rvq_expire_stack``).  Codebooks larger than the kernel's
    draw table keep the stage-by-stage path."""
    cbs = [l._codebook for l in layers]
    thr = cbs[0].threshold_ema_dead_code
    if thr == 0:
        return
    if not RUN_DEAD_EXPIRY:
        # core_vq.py:226 replaces dead rows of `embed`, core_vq.py:235 then overwrites EVERY row of `embed` with
        # embed_avg / smoothed cluster_size in the same call, and nothing in between reads `embed`: inside
        # a training forward the expiry cannot change any buffer.  Its launches are therefore skipped; the device generator
        # still advances as if the draw had happened.  (EuclideanCodebook.expire_codes_ called on its own does replace rows.)
        _rng_stream(x.device, len(cbs))
        return
    if 2 * cbs[0].codebook_size <= 4096:
        seed, off = _rng_stream(x.device, len(cbs))
        ops.expire_stack(pk, x, codes, stage0, [cb.cluster_size for cb in cbs], [cb.embed for cb in cbs],
                         float(thr), seed, off, flags)
        for cb in cbs:
            cb._tables_changed()
        return
    flags_host = torch.stack([(cb.cluster_size < thr).any() for cb in cbs]).tolist()
    b, _, t = x.shape
    for i, (cb, fire) in enumerate(zip(cbs, flags_host)):
        if not fire:
            continue
        sel = sample_indices(b * t, cb.codebook_size, x.device).contiguous()
        ops.expire_codes(pk, x, codes, stage0, i, sel, cb.cluster_size, float(thr), cb.embed, flags)
        cb._tables_changed()


def _stack_forward(layers: tp.Sequence[VectorQuantization], x: torch.Tensor, n_q: int, training: bool,
                   stack_pack: tp.Optional[tp.Callable[[], ops.CodebookPack]] = None, out_bdt: bool = False):
    """Forward of a residual stack of plain (un-projected) layers: core_vq.py:337-355 over
    :301-324 over :212-237.  Steady state is one fused launch for all ``n_q`` stages; a stage
    whose codebook still needs its k-means init forces that step to run stage by stage
    (SURVEY.md 3.4-9).  Returns ``(quantized [B, D, T], codes [n_q, B, T], losses [n_q, 1])``."""
    L.require_cuda_f32(x, "x")
    if x.dim() != 3:
        raise RuntimeError(f"expected x of shape [B, D, T], got {tuple(x.shape)}")
    layers = list(layers[:n_q])
    n_q = len(layers)
    cbs = [l._codebook for l in layers]
    cw = float(layers[0].commitment_weight)
    xd = x.detach()
    b, d, t = xd.shape
    flags = L.FLAG_STE if training else 0
    if training:
        _warn_issue25()

    if all(cb._is_inited() for cb in cbs):
        pk = stack_pack() if stack_pack is not None else ops.pack([cb.embed for cb in cbs])
        if training:
            # The sum of the straight-through values telescopes: sum_i ste_i = x - r_final (each stage subtracts what it
            # adds, core_vq.py:348-349), so the fused search hands back its final residual and the quantized sum is one
            # elementwise pass instead of a second walk over all stages' rows; fp32 rounding differs from the reference's
            # running sum by a few ulp (<< the 1e-5 bar).  Eval keeps the ordered sum of gathered rows (bit-exact).
            # The EMA statistics (core_vq.py:227-228) come out of the same launch: the search adds each stage's input
            # residual to its code's row while the row is on chip (rvq_encode_train).
            stats = ops.ema_stats_buffer(n_q, pk.K, pk.D, xd.device)
            codes, _, sqerr, res = ops.encode(pk, xd, 0, n_q, want_sqerr=cw > 0, want_residual=True, flags=flags,
                                              ema_stats_out=stats)
            quant = (xd - res.permute(0, 2, 1)) if out_bdt else (xd.permute(0, 2, 1) - res)
            if out_bdt and not quant.is_contiguous():
                quant = quant.contiguous()
            with torch.no_grad():
                _expire_stack(layers, pk, xd, codes, 0, flags)
                _update_stack(cbs, pk, xd, codes, 0, flags, stats=stats)
        else:
            codes, quant, sqerr, _ = ops.encode(pk, xd, 0, n_q, want_quantized=True, flags=flags, out_bdt=out_bdt)
    else:
        # first step(s): stage i's init needs stage i-1's fresh quantisation -> sequential
        quant = torch.zeros((b, t, d), dtype=torch.float32, device=x.device)
        res = xd
        code_list, sq_list, snap = [], [], []
        for cb in cbs:
            if not cb._is_inited():
                cb.init_embed_(res.permute(0, 2, 1).reshape(b * t, d))
            pk_i = cb._pack()
            snap.append(cb.embed.clone() if training else cb.embed)
            c_i, _, sq_i, res_next = ops.encode(pk_i, res, 0, 1, quantized_accum=quant,
                                                want_sqerr=training and cw > 0, want_residual=True, flags=flags)
            if training:
                with torch.no_grad():
                    cb.expire_codes_(res.permute(0, 2, 1).reshape(b * t, d))
                    _update_stack([cb], pk_i, res, c_i, 0, flags)
            res = res_next.permute(0, 2, 1)
            code_list.append(c_i)
            sq_list.append(sq_i)
        codes = torch.cat(code_list, 0)
        sqerr = torch.cat(sq_list, 0) if sq_list[0] is not None else None
        pk = ops.pack(snap) if (training and x.requires_grad) else None
        if out_bdt:
            quant = quant.permute(0, 2, 1).contiguous()

    if training and cw > 0:
        losses = (sqerr / float(b * t * d)).to(torch.float32).mul_(cw).view(n_q, 1)
    else:
        losses = torch.zeros((n_q, 1), dtype=torch.float32, device=x.device)
    # [B, D, T]: a permuted view of the [B, T, D] buffer like the reference's rearrange (core_vq.py:322), or -- with
    # contiguous_outputs -- the contiguous tensor the kernel wrote directly
    quantized = quant if out_bdt else quant.permute(0, 2, 1)
    if training:
        if x.requires_grad and torch.is_grad_enabled():
            quantized, losses = _AttachGrad.apply(x, quantized, losses, codes, pk, n_q, cw, flags)
        else:
            losses.requires_grad_(torch.is_grad_enabled())
    return quantized, codes, losses


class ResidualVectorQuantization(nn.Module):
    """Residual vector quantization (core_vq.py:327-375): Algorithm 1 of SoundStream, with the
    whole ``n_q``-stage loop fused into one kernel launch per call."""

    def __init__(self, *, num_quantizers, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([VectorQuantization(**kwargs) for _ in range(num_quantizers)])
        self._pack_cache: tp.Optional[tp.Tuple[tp.Any, ops.CodebookPack]] = None
        # False: `quantized` / decode outputs are permuted views of a [B, T, D] buffer, strides as in the reference
        # (core_vq.py:298, :322).  True: contiguous [B, D, T] tensors written directly by the kernels -- what SEANet's decoder
        # convolution consumes (modules/seanet.py:193-195) without a re-layout copy.
        self.contiguous_outputs = False

    def _require_plain(self, n: int) -> None:
        for l in self.layers[:n]:
            l._require_plain()

    def _stack_pack(self) -> ops.CodebookPack:
        """Search image of ALL stages (a prefix serves any ``n_q``), cached until a table changes."""
        key = tuple(l._codebook._key() for l in self.layers)
        if self._pack_cache is None or self._pack_cache[0] != key:
            self._pack_cache = (key, ops.pack([l._codebook.embed for l in self.layers]))
        return self._pack_cache[1]

    def invalidate(self) -> None:
        for l in self.layers:
            l._codebook.invalidate()
        self._pack_cache = None

    def forward(self, x: torch.Tensor, n_q: tp.Optional[int] = None):
        """core_vq.py:337-355."""
        n_q = n_q or len(self.layers)
        n_q = min(n_q, len(self.layers))           # the reference's slice caps silently (:346)
        self._require_plain(n_q)
        return _stack_forward(self.layers, x, n_q, self.training, self._stack_pack, out_bdt=self.contiguous_outputs)

    def encode(self, x: torch.Tensor, n_q: tp.Optional[int] = None, layout: str = "kbt") -> torch.Tensor:
        """core_vq.py:357-367: ``[B, D, T]`` fp32 -> ``[n_q, B, T]`` int64.  ``layout="bkt"`` returns the contiguous
        ``[B, n_q, T]`` tensor that model.py:166 obtains with ``codes.transpose(0, 1)`` (same values, written directly)."""
        if layout not in ("kbt", "bkt"):
            raise ValueError(f"layout must be 'kbt' or 'bkt', got {layout!r}")
        n_q = n_q or len(self.layers)
        n_q = min(n_q, len(self.layers))
        self._require_plain(n_q)
        L.require_cuda_f32(x, "x")
        codes, _, _, _ = ops.encode(self._stack_pack(), x.detach(), 0, n_q, codes_bkt=layout == "bkt")
        return codes

    def decode(self, q_indices: torch.Tensor) -> torch.Tensor:
        """core_vq.py:369-375: ``[n_q, B, T]`` int64 (any strides, any stage prefix) ->
        ``[B, D, T]`` fp32, stages summed in order."""
        n = int(q_indices.shape[0])
        self._require_plain(n)
        if self.contiguous_outputs:
            return ops.decode(self._stack_pack(), q_indices, out_bdt=True)
        return ops.decode(self._stack_pack(), q_indices).permute(0, 2, 1)
