"""The callers' side of the hot path: code / latent layouts handed to and from SEANet
(model.py:165-166, :188-189) and the 48 kHz model's one-second segment loop (model.py:141-145) in one launch.  Every
variant is checked against the CPU oracle, called the way the reference calls the quantizer."""
import pytest
import torch

from oracle import cases as C
from oracle import rvq_oracle as O

from helpers import assert_codes_match, build_module, module_states

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3, 203, 12), (2, 37, 32), (1, 1, 2)], ids=["b3_t203_nq12", "b2_t37_nq32", "b1_t1_nq2"])
@pytest.mark.parametrize("exact", [False, True], ids=["tc", "simt"])
def test_codes_bkt_and_contiguous_outputs(shape, exact):
    from encodec_pytorch_b200 import _lib as L, _ops as ops
    b, t, n_q = shape
    case = C.Case("lay", b, 128, t, 1024, n_q, 75, None, 505, 12)
    q = build_module(case).eval()
    states = module_states(q)
    x = C.latents(b, 128, t, case.x_seed)
    xg = x.cuda()
    pk = q.vq._stack_pack()
    fl = L.FLAG_FORCE_EXACT if exact else 0
    kbt, quant, _, _ = ops.encode(pk, xg, 0, n_q, want_quantized=True, flags=fl)
    bkt, quant_bdt, _, _ = ops.encode(pk, xg, 0, n_q, want_quantized=True, flags=fl, codes_bkt=True, out_bdt=True)
    assert bkt.shape == (b, n_q, t) and bkt.is_contiguous() and quant_bdt.shape == (b, 128, t) and quant_bdt.is_contiguous()
    assert torch.equal(bkt, kbt.transpose(0, 1))                       # model.py:166: codes.transpose(0, 1)
    assert torch.equal(quant_bdt, quant.permute(0, 2, 1))              # same sums, written [B, D, T]
    assert_codes_match(states, x, kbt, O.rvq_encode(states, x).numpy())
    # decode: [B, K, T] frames come back as transposed views (model.py:188), output contiguous [B, D, T]
    dec = ops.decode(pk, bkt.transpose(0, 1), out_bdt=True)
    assert dec.is_contiguous() and torch.equal(dec.cpu(), O.rvq_decode(states, kbt.cpu()).contiguous())
    # accumulate across stage segments in the [B, D, T] layout
    half = max(1, n_q // 2)
    c1, q1, _, r1 = ops.encode(pk, xg, 0, half, want_quantized=True, want_residual=True, flags=fl, out_bdt=True)
    if half < n_q:
        c2, q2, _, _ = ops.encode(pk, r1.permute(0, 2, 1), half, n_q - half, quantized_accum=q1, flags=fl, out_bdt=True)
        assert torch.equal(torch.cat([c1, c2], 0), kbt) and torch.equal(q2, quant_bdt)


def test_module_options_match_the_reference_layouts():
    case = C.Case("opt", 2, 128, 150, 1024, 16, 150, 24.0, 606, 13)
    q = build_module(case).eval()
    x = C.latents(2, 128, 150, case.x_seed).cuda()
    with torch.no_grad():
        codes = q.encode(x, 150, 24.0)
        bkt = q.encode(x, 150, 24.0, layout="bkt")
        assert bkt.is_contiguous() and torch.equal(bkt, codes.transpose(0, 1))
        ref_dec = q.decode(codes)
        ref_fwd = q(x, 150, 24.0).quantized
        assert not ref_dec.is_contiguous()                                 # the reference's permuted view (core_vq.py:298)
        q.contiguous_outputs = True
        dec = q.decode(bkt.transpose(0, 1))
        fwd = q(x, 150, 24.0).quantized
        assert dec.is_contiguous() and fwd.is_contiguous()
        assert torch.equal(dec, ref_dec) and torch.equal(fwd, ref_fwd)
    q.train()
    xg = x.clone().requires_grad_(True)
    with pytest.warns(UserWarning):
        res = q(xg, 150, 24.0)
    assert res.quantized.is_contiguous() and res.quantized.shape == x.shape
    (res.quantized.sum() + res.penalty).backward()
    assert torch.isfinite(xg.grad).all()
    with pytest.raises(ValueError):
        q.encode(x, 150, 24.0, layout="tkb")


def test_segment_batch_equals_the_reference_loop():
    """model.py:141-145 encodes a 30 s clip of the 48 kHz model as 31 one-second segments (150 frames each, the last one
    shorter), one quantizer call per segment; encode_segments / decode_segments do the same with one launch.  Checked against
    the oracle called once per segment, as the reference loop does."""
    case = C.Case("seg", 4, 128, 150, 1024, 16, 150, 24.0, 707, 14)
    q = build_module(case).eval()
    states = module_states(q)
    lens = [150] * 30 + [45]
    segs = [C.latents(4, 128, n, 900 + i) for i, n in enumerate(lens)]
    with torch.no_grad():
        got = q.encode_segments([s.cuda() for s in segs], 150, 24.0)
        got_bkt = q.encode_segments([s.cuda() for s in segs], 150, 24.0, layout="bkt")
        one_by_one = [q.encode(s.cuda(), 150, 24.0) for s in segs]
    assert [tuple(c.shape) for c in got] == [(16, 4, n) for n in lens]
    for c, cb, c1, s in zip(got, got_bkt, one_by_one, segs):
        assert torch.equal(c, c1) and torch.equal(cb, c.transpose(0, 1))    # frames are independent: same codes either way
        assert_codes_match(states, s, c, O.rvq_encode(states, s).numpy())
    with torch.no_grad():
        dec = q.decode_segments([cb.transpose(0, 1) for cb in got_bkt])       # [K, B, T] views of [B, K, T] frames (model.py:188)
    for d, c in zip(dec, got):
        assert torch.equal(d.cpu(), O.rvq_decode(states, c.cpu()))
    assert q.encode_segments([], 150) == []
