"""Oracle parity at the sizes BASELINE.json names (configs[1..4]): the CUDA path through the public module API against
the CPU oracle (``oracle/rvq_oracle.py``, pinned to the reference by ``tests/test_oracle_golden.py``) on the same seeded
inputs.  The oracle needs a few seconds per case on the GPU box's host cores.

Codes: equal to the oracle's free-running codes, or -- where an fp32 near-tie moved a code -- equal stage by stage to the
oracle's search on the candidate's own residual chain, every difference being a documented near-tie (fp64 distance gap
< 1e-6 relative; BASELINE.json north_star).  Floating-point outputs: 1e-5 relative (to the tensor's scale where sums of
thousands of rows are compared)."""
import warnings

import pytest
import torch

from oracle import cases as C
from oracle import rvq_oracle as O

from helpers import assert_codes_match, build_module, module_states

pytestmark = pytest.mark.gpu

CFG2 = C.Case("cfg2_b64_t750_nq32", 64, 128, 750, 1024, 32, 75, 24.0, 1234, 0)
CFG4 = C.Case("cfg4_b32_t4500_nq16", 32, 128, 4500, 1024, 16, 150, 24.0, 1234, 0)


def _encode_decode_against_oracle(case, max_near_frac=2e-3):
    q = build_module(case).eval()
    states = module_states(q)
    x = C.latents(case.b, case.d, case.t, case.x_seed)
    with torch.no_grad():
        codes = q.encode(x.cuda(), case.frame_rate, case.bandwidth)
    want = O.rvq_encode(states, x, n_q=codes.shape[0])
    st = assert_codes_match(states, x, codes, want.numpy(), max_near_frac=max_near_frac)
    # decode of the kernel's own codes: bit-exact against the oracle's ordered sum of gathered rows
    with torch.no_grad():
        dec = q.decode(codes)
    assert torch.equal(dec.cpu(), O.rvq_decode(states, codes.cpu()))
    return st


def test_cfg2_encode_decode_full_size():
    """configs[1]: 24 kHz 24 kbps, [64, 128, 750], n_q = 32 (48 000 frames x 32 stages)."""
    st = _encode_decode_against_oracle(CFG2)
    assert st["pairs"] == 64 * 750 * 32


def test_cfg4_encode_decode_full_size():
    """configs[3]: 48 kHz model, [32, 128, 4500], n_q = 16 at 150 frames/s (144 000 frames x 16 stages)."""
    st = _encode_decode_against_oracle(CFG4)
    assert st["pairs"] == 32 * 4500 * 16


@pytest.mark.parametrize("n_q", [2, 8])
def test_cfg5_point_1e5_frames(n_q):
    """configs[4]: one point of the bulk sweep, 1e5 frames ([134, 128, 750]) at n_q = 2 and 8."""
    case = C.Case(f"cfg5_1e5_nq{n_q}", 134, 128, 750, 1024, n_q, 75, None, 4242, 0)
    st = _encode_decode_against_oracle(case)
    assert st["pairs"] == 134 * 750 * n_q


def _teacher_forced_training_expectation(pre_states, x, codes, decay=0.99, eps=1e-5, cw=1.0):
    """What core_vq.py:337-355 / :212-237 produce for one training step when the searches return ``codes``: quantized sum
    of the straight-through values, commitment losses, and the buffers after the EMA update (the expiry of :165-175 draws
    random rows into ``embed``, which :235 overwrites in the same step -- a net no-op on every buffer).
    Built from the oracle's own primitives on the candidate's residual chain."""
    n_q, b, t = codes.shape
    states = [{k: v.clone() for k, v in s.items()} for s in pre_states[:n_q]]
    res = O.frames_of(x)
    total = torch.zeros_like(res)
    losses = []
    for i, st in enumerate(states):
        idx = codes[i].reshape(-1).to(torch.long)
        qv = O.lookup(idx, st["embed"])
        qv = res + (qv - res)                                   # core_vq.py:309
        losses.append(torch.nn.functional.mse_loss(qv, res) * cw)
        O.ema_update(st, res, idx, decay, eps)
        res = res - qv                                          # core_vq.py:348
        total = total + qv
    return O.unframe(total, b, t), torch.stack(losses), states


def test_cfg3_training_step_full_size():
    """configs[2] on one rank: a training forward at [64, 128, 750] x 32 from random-init codebooks (every stage's expiry
    fires in this step): codes, penalty, quantized, cluster_size / embed_avg / embed after the EMA update."""
    case = CFG2
    q = build_module(case).train()
    pre = module_states(q)
    x = C.latents(case.b, case.d, case.t, case.x_seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad():
            res = q(x.cuda(), case.frame_rate, case.bandwidth)
    codes = res.codes.cpu()
    st = O.compare_codes_teacher_forced(pre, x, codes)
    assert st["bad"] == 0 and st["near_tie"] <= 2e-3 * st["pairs"], st
    want_q, want_l, post = _teacher_forced_training_expectation(pre, x, codes)
    assert res.penalty.item() == pytest.approx(want_l.mean().item(), rel=1e-5)
    torch.testing.assert_close(res.quantized.cpu(), want_q, rtol=1e-5, atol=1e-5 * float(want_q.abs().max()))
    for i in (0, 1, 7, 15, 31):
        cb = q.vq.layers[i]._codebook
        torch.testing.assert_close(cb.cluster_size.cpu(), post[i]["cluster_size"], rtol=1e-5, atol=1e-6)
        ea, em = post[i]["embed_avg"], post[i]["embed"]
        # float atomics sum ~47 residual rows per code in arbitrary order: 1e-5 relative to the tensor's scale
        torch.testing.assert_close(cb.embed_avg.cpu()[::7], ea[::7], rtol=1e-5, atol=1e-5 * float(ea.abs().max()))
        torch.testing.assert_close(cb.embed.cpu()[::7], em[::7], rtol=2e-5, atol=1e-5 * float(em.abs().max()))


def test_cfg2_trained_like_stack():
    """SURVEY.md 8(d)'s "trained-like" run at configs[1] size: k-means init + 25 EMA steps fit the codebooks to the latents
    (residual norms decay over the stages, ~13 % of the frame-stages need the exact re-score), then an eval encode of fresh
    latents is checked against the oracle holding the same fitted tables."""
    torch.manual_seed(0)
    import encodec_pytorch_b200 as E
    n_q = 32
    q = E.ResidualVectorQuantizer(dimension=128, n_q=n_q, bins=1024, kmeans_init=True, kmeans_iters=10).cuda().train()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad():
            for i in range(26):
                q(C.latents(64, 128, 750, 500 + i).cuda(), 75, 24.0)
    q.eval()
    states = module_states(q)
    x = C.latents(64, 128, 750, 900)
    # the three settings of the score-error bound (rvq_pack_bound_mode): per-stage everywhere, per-code everywhere, chosen
    # per stage.  All must produce the oracle's codes; on fitted tables the per-code bound certifies far more frames.
    from encodec_pytorch_b200 import _ops as ops
    share = {}
    for mode in (2, 1, 0):
        with ops.pack_bound_mode(mode):
            q.vq.invalidate()
            with ops.search_counters(torch.device("cuda", 0)) as counters, torch.no_grad():
                codes = q.encode(x.cuda(), 75, 24.0)
        c = counters.read()
        share[mode] = c["certified"] / c["searched"]
        st = O.compare_codes_teacher_forced(states, x, codes.cpu())
        assert st["bad"] == 0 and st["near_tie"] <= 2e-3 * st["pairs"], (mode, st)
    assert share[1] > share[2] + 0.03 and share[0] >= share[1] - 0.01, share
    with torch.no_grad():
        dec = q.decode(codes)
    assert torch.equal(dec.cpu(), O.rvq_decode(states, codes.cpu()))
    # the stack is really fitted: the residual after all stages is well below the input
    assert float((x - dec.cpu()).norm() / x.norm()) < 0.7


@pytest.mark.parametrize("mode", [1, 2])
def test_bound_modes_random_init(mode):
    """Freshly initialised tables (uniform norms) under the per-code and the per-stage bound: same codes as the oracle."""
    from encodec_pytorch_b200 import _ops as ops
    case = C.Case(f"bound_mode{mode}", 16, 128, 750, 1024, 8, 75, None, 77, 0)
    with ops.pack_bound_mode(mode):
        _encode_decode_against_oracle(case)
