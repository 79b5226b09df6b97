"""Pin the CPU oracle (oracle/rvq_oracle.py) against fixtures produced by the
real reference (oracle/gen_golden.py; SURVEY.md 8(c): the reference ships no
golden vectors for this path, so these are outputs of the reference run in the
build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import cases as C
from oracle import rvq_oracle as O


def load(golden_dir, kind, case):
    return np.load(os.path.join(golden_dir, f"{kind}_{case.name}.npz"))


def check_summary(t, g, prefix, stride=97, rtol=0.0, atol=0.0):
    c = t.detach().contiguous()
    assert list(c.shape) == list(g[prefix + "_shape"])
    if rtol == 0.0 and atol == 0.0 and C.sha(c) == str(g[prefix + "_sha"]):
        return
    sub = c.reshape(-1)[::stride].numpy()
    np.testing.assert_allclose(sub, g[prefix + "_sub"], rtol=max(rtol, 1e-6), atol=max(atol, 1e-6))
    assert abs(c.double().abs().sum().item() - float(g[prefix + "_abs64"])) <= 1e-5 * float(g[prefix + "_abs64"]) + 1e-6


def assert_codes_match(states, x, got, want):
    """Exact on the generating host; on another CPU only fp32 near-ties may differ."""
    want = torch.from_numpy(want.astype(np.int64))
    if torch.equal(got, want):
        return
    st = O.compare_codes_teacher_forced(states, x, want)
    assert st["bad"] == 0, st


@pytest.mark.parametrize("case", C.ENCODE_CASES, ids=lambda c: c.name)
def test_encode_decode_forward_eval(golden_dir, case):
    g = load(golden_dir, "encode", case)
    states = C.codebooks(case.d, case.k, case.n_q, case.cb_seed)
    x = C.latents(case.b, case.d, case.t, case.x_seed, case.x_scale)
    assert C.sha(x) == str(g["x_sha"])
    assert C.sha(states[0]["embed"]) == str(g["embed0_sha"])
    n_q = O.num_quantizers_for_bandwidth(case.n_q, case.k, case.frame_rate, case.bandwidth)
    codes = O.rvq_encode(states, x, n_q)
    assert codes.shape == tuple(g["codes"].shape)
    assert_codes_match(states, x, codes, g["codes"])
    want = torch.from_numpy(g["codes"].astype(np.int64))
    dec = O.rvq_decode(states, want)
    # strides of size-1 dims are arbitrary; the rest must be the reference's permuted-[B,T,D] view
    assert [s_ for s_, n in zip(dec.stride(), dec.shape) if n > 1] == \
        [int(s_) for s_, n in zip(g["decode_strides"], dec.shape) if n > 1]
    check_summary(dec, g, "decode")
    check_summary(O.rvq_decode(states, want[: max(1, want.shape[0] // 2)]), g, "decode_prefix")
    res = O.quantizer_forward(states, x, case.frame_rate, case.bandwidth, case.k, training=False)
    check_summary(res["quantized"], g, "fwd_quantized")
    assert res["bandwidth"].item() == pytest.approx(float(g["fwd_bandwidth"]), rel=0, abs=0)
    assert res["penalty"].item() == 0.0 == float(g["fwd_penalty"])
    assert res["bandwidth"].dim() == 0 and res["penalty"].dim() == 0


def test_cfg1_fingerprint(golden_dir):
    """SURVEY.md 3.4-12 known-answer record."""
    g = load(golden_dir, "encode", C.ENCODE_CASES[0])
    assert int(g["codes"].astype(np.int64).sum()) == 12272372
    assert float(g["decode_sum64"]) == pytest.approx(776.2944878875569, rel=1e-12)


@pytest.mark.parametrize("case", C.TRAIN_CASES, ids=lambda c: c.name)
def test_training_forward_ema_and_grad(golden_dir, case):
    g = load(golden_dir, "train", case)
    states = C.codebooks(case.d, case.k, case.n_q, case.cb_seed)
    n_q = O.num_quantizers_for_bandwidth(case.n_q, case.k, case.frame_rate, case.bandwidth)
    torch.manual_seed(1000 + case.cb_seed)
    for s in range(3):
        x = C.latents(case.b, case.d, case.t, case.x_seed + s, case.x_scale)
        w = C.latents(case.b, case.d, case.t, 5000 + s)
        assert C.sha(x) == str(g[f"s{s}_x_sha"])
        pre = [{k: v.clone() for k, v in st.items()} for st in states]
        grad = O.rvq_forward_grad(pre, x, n_q, w, torch.full((n_q, 1), 3.0 / n_q))
        res = O.quantizer_forward(states, x, case.frame_rate, case.bandwidth, case.k, training=True)
        assert_codes_match(pre, x, res["codes"], g[f"s{s}_codes"])
        assert res["penalty"].item() == pytest.approx(float(g[f"s{s}_penalty"]), rel=1e-6)
        check_summary(res["quantized"], g, f"s{s}_quantized")
        check_summary(grad, g, f"s{s}_grad", rtol=1e-5, atol=1e-7)
    for i, st in enumerate(states):
        np.testing.assert_allclose(st["cluster_size"].numpy(), g[f"L{i}_cluster_size"], rtol=1e-6, atol=1e-7)
        check_summary(st["embed"], g, f"L{i}_embed", stride=31)
        check_summary(st["embed_avg"], g, f"L{i}_embed_avg", stride=31)
        assert st["inited"].item() == float(g[f"L{i}_inited"][0])


@pytest.mark.parametrize("case", C.KMEANS_CASES, ids=lambda c: c.name)
def test_kmeans_init_first_forward(golden_dir, case):
    g = load(golden_dir, "kmeans", case)
    states = C.codebooks(case.d, case.k, case.n_q, case.cb_seed, kmeans_init=True)
    x = C.latents(case.b, case.d, case.t, case.x_seed, case.x_scale)
    assert C.sha(x) == str(g["x_sha"])
    # un-inited kmeans tables are all-zero: encode() never initialises
    assert int(O.rvq_encode(states, x).max()) == 0 == int(g["uninited_codes_max"])
    torch.manual_seed(2000 + case.cb_seed)
    res = O.quantizer_forward(states, x, case.frame_rate, case.bandwidth, case.k, training=True,
                              kmeans_iters=int(g["iters"]))
    assert torch.equal(res["codes"], torch.from_numpy(g["codes"].astype(np.int64)))
    assert res["penalty"].item() == pytest.approx(float(g["penalty"]), rel=1e-6)
    check_summary(res["quantized"], g, "quantized")
    for i, st in enumerate(states):
        np.testing.assert_allclose(st["cluster_size"].numpy(), g[f"L{i}_cluster_size"], rtol=1e-6, atol=1e-6)
        check_summary(st["embed"], g, f"L{i}_embed", stride=31)
        assert st["inited"].item() == 1.0


def test_reference_direct_when_present():
    """In the build container the real reference is importable: compare directly."""
    ref = os.environ.get("RVQ_REFERENCE_DIR", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "quantization")):
        pytest.skip("reference tree not present (GPU box)")
    import sys
    sys.path.insert(0, ref)
    try:
        import quantization as qt
    finally:
        sys.path.remove(ref)
    torch.manual_seed(11)
    q = qt.ResidualVectorQuantizer(dimension=32, n_q=5, bins=128, kmeans_init=False).eval()
    states = O.states_from_module(q)
    x = C.latents(3, 32, 41, 8)
    with torch.no_grad():
        want = q.encode(x, 75, None)
        assert torch.equal(O.rvq_encode(states, x), want)
        assert torch.equal(O.rvq_decode(states, want), q.decode(want))
