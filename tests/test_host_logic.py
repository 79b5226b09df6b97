"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
host mirror keeps the reference's API surface, and compute entry points fail loudly without a
B200 (no fallback).  No GPU needed."""
import ctypes
import math
import os
import re

import pytest
import torch

from oracle import rvq_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rvq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rvq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from encodec_pytorch_b200 import _lib as L
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rvq_b200.h but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(L.SIGNATURES) == names
    assert lib.rvq_version() == 1
    assert lib.rvq_pack_bytes(32, 1024, 128) > 32 * 1024 * 128 * 4
    assert lib.rvq_pack_bytes(0, 1024, 128) == 256
    assert isinstance(lib.rvq_last_error(), bytes)


def test_pack_bound_mode_switch():
    """rvq_pack_bound_mode is host state only: it hands back the previous mode, clamps garbage to the default, and the
    context manager restores what it found."""
    from encodec_pytorch_b200 import _lib as L, _ops as ops
    lib = L.load()
    prev = lib.rvq_pack_bound_mode(0)
    try:
        assert lib.rvq_pack_bound_mode(1) == 0 and lib.rvq_pack_bound_mode(2) == 1
        assert lib.rvq_pack_bound_mode(7) == 2 and lib.rvq_pack_bound_mode(-3) == 0      # out of range -> default
        assert lib.rvq_pack_bound_mode(0) == 0
        with ops.pack_bound_mode(2):
            with ops.pack_bound_mode(1):
                assert lib.rvq_pack_bound_mode(1) == 1
            assert lib.rvq_pack_bound_mode(2) == 2
        assert lib.rvq_pack_bound_mode(0) == 0
    finally:
        lib.rvq_pack_bound_mode(prev)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_compute_entry_points_refuse_without_device():
    from encodec_pytorch_b200 import _lib as L
    lib = L.load()
    assert lib.rvq_device_ok() == -2          # RVQ_ENODEV
    buf = (ctypes.c_float * 4)()
    rc = lib.rvq_decode(buf, 4, 4, buf, 1, 1, 1, 1, 1, 1, buf, None)
    assert rc == -2 and b"no CPU fallback" in lib.rvq_last_error() or b"fallback" in lib.rvq_last_error()


def test_api_surface_and_defaults():
    import encodec_pytorch_b200 as E
    from encodec_pytorch_b200.quantization import core_vq, vq
    q = E.ResidualVectorQuantizer()
    assert (q.dimension, q.n_q, q.bins, q.decay, q.kmeans_init, q.kmeans_iters, q.threshold_ema_dead_code) == \
        (256, 8, 1024, 0.99, True, 50, 2)
    assert isinstance(q.vq, core_vq.ResidualVectorQuantization) and len(q.vq.layers) == 8
    layer = q.vq.layers[0]
    assert isinstance(layer, core_vq.VectorQuantization) and isinstance(layer._codebook, core_vq.EuclideanCodebook)
    assert isinstance(layer.project_in, torch.nn.Identity) and isinstance(layer.project_out, torch.nn.Identity)
    cb = layer._codebook
    assert cb.inited.tolist() == [0.0] and float(cb.embed.abs().sum()) == 0.0      # kmeans_init=True -> zeros
    assert len(list(q.parameters())) == 0                                          # buffers only
    assert [n for n, _ in cb.named_buffers()] == ["inited", "cluster_size", "embed", "embed_avg"]
    for name in ("init_embed_", "replace_", "expire_codes_", "preprocess", "quantize", "postprocess_emb",
                 "dequantize", "encode", "decode", "forward"):
        assert callable(getattr(cb, name))
    for name in ("default", "ema_inplace", "laplace_smoothing", "uniform_init", "sample_vectors", "kmeans"):
        assert callable(getattr(core_vq, name))
    r = vq.QuantizedResult(torch.zeros(1), torch.zeros(1), torch.zeros(()))
    assert r.penalty is None and r.metrics == {}
    proj = core_vq.VectorQuantization(dim=32, codebook_size=16, codebook_dim=8)
    assert isinstance(proj.project_in, torch.nn.Linear) and proj._codebook.embed.shape == (16, 8)


def test_constructor_draws_the_reference_tables():
    """Same RNG consumption as the reference constructor: kaiming-uniform per stage, in order."""
    import encodec_pytorch_b200 as E
    torch.manual_seed(0)
    q = E.ResidualVectorQuantizer(dimension=128, n_q=3, bins=1024, kmeans_init=False)
    torch.manual_seed(0)
    states = O.new_rvq_states(128, 1024, 3, False)
    for layer, st in zip(q.vq.layers, states):
        assert torch.equal(layer._codebook.embed, st["embed"]) and torch.equal(layer._codebook.embed_avg, st["embed"])
        assert layer._codebook.inited.tolist() == [1.0]
    # SURVEY.md 3.4-12 fingerprint of the first table
    assert q.vq.layers[0]._codebook.embed[0, :3].tolist() == pytest.approx([-0.00162094, 0.11614344, -0.17819449], abs=1e-7)


@pytest.mark.parametrize("frame_rate,bins,n_total", [(75, 1024, 32), (150, 1024, 16), (50, 2048, 8)])
def test_bandwidth_to_stage_count(frame_rate, bins, n_total):
    import encodec_pytorch_b200 as E
    q = E.ResidualVectorQuantizer(dimension=8, n_q=n_total, bins=bins, kmeans_init=False)
    assert q.get_bandwidth_per_quantizer(frame_rate) == O.bandwidth_per_quantizer(bins, frame_rate)
    for bw in (None, 0, 0.0, 0.1, 1.5, 3, 6.0, 12, 24.0, 48.0, -1.0):
        assert q.get_num_quantizers_for_bandwidth(frame_rate, bw) == \
            O.num_quantizers_for_bandwidth(n_total, bins, frame_rate, bw)
    if frame_rate == 75:
        assert [q.get_num_quantizers_for_bandwidth(75, b) for b in (1.5, 3, 6, 12, 24)] == [2, 4, 8, 16, 32]
    if frame_rate == 150:
        assert [q.get_num_quantizers_for_bandwidth(150, b) for b in (3, 6, 12, 24)] == [2, 4, 8, 16]


def test_cpu_tensors_raise_not_fall_back():
    import encodec_pytorch_b200 as E
    q = E.ResidualVectorQuantizer(dimension=128, n_q=2, bins=1024, kmeans_init=False).eval()
    x = torch.randn(1, 128, 4)
    for call in (lambda: q.encode(x, 75), lambda: q(x, 75), lambda: q.decode(torch.zeros(2, 1, 4, dtype=torch.long))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_missing_library_is_a_loud_error(monkeypatch):
    from encodec_pytorch_b200 import _lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        L.load()


def test_shard_frames_covers_batch():
    from encodec_pytorch_b200 import distrib
    for batch in (1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [distrib.shard_frames(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "encodec_pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "rvq_oracle" not in src, f


def test_projected_codebooks_raise_instead_of_falling_back():
    """codebook_dim != dim is never built by the EnCodec models; the B200 path raises rather than run a per-op loop."""
    import pytest
    import torch
    from encodec_pytorch_b200.quantization.core_vq import ResidualVectorQuantization, VectorQuantization
    vq = VectorQuantization(dim=8, codebook_size=4, codebook_dim=4, kmeans_init=False)
    for call in (lambda: vq.encode(torch.zeros(1, 8, 3)), lambda: vq.decode(torch.zeros(1, 3, dtype=torch.long)),
                 lambda: vq(torch.zeros(1, 8, 3))):
        with pytest.raises(RuntimeError, match="projected codebooks"):
            call()
    rvq = ResidualVectorQuantization(num_quantizers=2, dim=8, codebook_size=4, codebook_dim=4, kmeans_init=False)
    with pytest.raises(RuntimeError, match="projected codebooks"):
        rvq.encode(torch.zeros(1, 8, 3))
