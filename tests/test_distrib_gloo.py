"""World-size-2 checks of the RVQ path's collective plumbing (``encodec_pytorch_b200/distrib.py``) on CPU
with the ``gloo`` backend: frames sharded over ranks, ONE all-reduce of the packed EMA statistics, rank-0
broadcast after k-means init.  The per-rank arithmetic here is the oracle's (the kernels need a B200); what
is under test is that P ranks with all-reduced statistics reproduce one rank on the concatenated batch
 through the very helpers the CUDA path calls."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cases as C
from oracle import rvq_oracle as O

D, K, NQ, B, T = 16, 64, 3, 4, 60


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_stats(states, x, n_q):
    """counts [n_q,K] and embed_sum [n_q,K,D] of one shard, stage by stage with the PRE-update tables
    (core_vq.py:219, :227-228), packed like ``_ops.ema_stats``: counts first, then the sums."""
    flat = O.frames_of(x)
    counts, sums, codes = [], [], []
    r = flat
    for i in range(n_q):
        idx = O.nearest_code(r, states[i]["embed"])
        onehot = torch.nn.functional.one_hot(idx, K).to(r.dtype)
        counts.append(onehot.sum(0))
        sums.append((r.t() @ onehot).t())
        q = O.lookup(idx, states[i]["embed"])
        q = r + (q - r)                      # straight-through value of core_vq.py:309 feeds the residual (:348)
        r = r - q
        codes.append(idx)
    return torch.cat([torch.stack(counts).reshape(-1), torch.stack(sums).reshape(-1)]), torch.stack(codes)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from encodec_pytorch_b200 import distrib
        assert distrib.world_size() == world and distrib.rank() == rank and distrib.is_distributed()
        # --- rank-0 broadcast of freshly initialised buffers (distrib.py:55-68 semantics) ---
        states = C.codebooks(D, K, NQ, 5)
        if rank != 0:
            for st in states:
                for v in st.values():
                    v.zero_()
        bufs = [v for st in states for v in st.values()]
        distrib.broadcast_tensors(bufs, src=0)
        ref_states = C.codebooks(D, K, NQ, 5)
        for st, rs in zip(states, ref_states):
            for k in st:
                assert torch.equal(st[k], rs[k]), k
        # --- frame-sharded EMA step: local statistics, one all-reduce, identical update on every rank ---
        x = C.latents(B, D, T, 77)
        lo, hi = distrib.shard_frames(B)
        flat, codes = _shard_stats(states, x[lo:hi], NQ)
        distrib.all_reduce_stats(flat)
        counts = flat[: NQ * K].view(NQ, K)
        esum = flat[NQ * K:].view(NQ, K, D)
        decay, eps = 0.99, 1e-5
        for i, st in enumerate(states):
            st["cluster_size"].mul_(decay).add_(counts[i], alpha=1 - decay)
            st["embed_avg"].mul_(decay).add_(esum[i], alpha=1 - decay)
            cs = (st["cluster_size"] + eps) / (st["cluster_size"].sum() + K * eps) * st["cluster_size"].sum()
            st["embed"].copy_(st["embed_avg"] / cs.unsqueeze(1))
        torch.save({"states": states, "codes": codes, "range": (lo, hi)}, os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank_on_the_concatenated_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    # single process, whole batch, the reference's interleaved order (expiry disabled: threshold 0)
    states = C.codebooks(D, K, NQ, 5)
    x = C.latents(B, D, T, 77)
    ref = O.rvq_forward(states, x, NQ, training=True, decay=0.99, eps=1e-5, kmeans_iters=0, threshold=0)
    ref_codes = ref[1] if isinstance(ref, tuple) else ref["codes"]
    for r in range(world):
        lo, hi = got[r]["range"]
        assert torch.equal(got[r]["codes"].view(NQ, hi - lo, T), ref_codes[:, lo:hi])
        for i in range(NQ):
            for k in ("cluster_size", "embed_avg", "embed"):
                torch.testing.assert_close(got[r]["states"][i][k], states[i][k], rtol=1e-5, atol=1e-6)
    # both ranks hold identical buffers afterwards (that is what keeps expiry decisions in sync)
    for i in range(NQ):
        for k in ("cluster_size", "embed_avg", "embed"):
            assert torch.equal(got[0]["states"][i][k], got[1]["states"][i][k])


def test_shard_frames_is_a_partition():
    from encodec_pytorch_b200 import distrib
    for b in (1, 2, 7, 64):
        for w in (1, 2, 4, 8):
            spans = [distrib.shard_frames(b, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(a[1] == c[0] for a, c in zip(spans, spans[1:]))
