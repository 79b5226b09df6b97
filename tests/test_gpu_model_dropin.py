"""Drop-in check against the reference's own callers (BASELINE.json north_star: "model.py's EncodecModel, compress.py
and train_multi_gpu.py use it as a drop-in").  The UNMODIFIED reference is imported from the git-ignored ``baseline/_ref``
(``scripts/install_reference.py``); its ``EncodecModel`` is built twice, once around the reference quantizer and once with
the one-line swap of INTEGRATION.md (``model.qt = encodec_pytorch_b200.quantization``), both holding the same state_dict:

  * ``model.encode`` (model.py:141-166): same codes, ``model.decode`` (:168-192): same waveform, 24 kHz and 48 kHz models;
  * ``.ecdc`` streams: ``encodec_pytorch_b200.compress.compress_to_file`` (one ``pack_frame`` launch per segment) writes the
    bytes the reference's ``compress.compress_to_file`` (compress.py:30-92, one ``BitPacker.push`` per value) writes;
  * training: ``model.train(); out, loss_w, _ = model(x); loss_w.backward()`` under ``DistributedDataParallel`` with
    ``broadcast_buffers=False`` (train_multi_gpu.py:61, :94, :318), one rank per visible GPU (up to 2).
"""
import io
import os
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    import install_reference as IR
    if not os.path.isdir(os.path.join(IR.DEST, "quantization")):
        if os.path.isdir("/root/reference"):
            IR.install()
        else:
            pytest.skip("baseline/_ref is not populated (run scripts/install_reference.py where the reference exists)")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return IR.import_reference()


def _build_pair(ref, which):
    """The reference model twice: around its own quantizer and around ours, same weights and codebooks."""
    import encodec_pytorch_b200.quantization as our_qt
    factory = {"24khz": ref.model.EncodecModel.encodec_model_24khz, "48khz": ref.model.EncodecModel.encodec_model_48khz}[which]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(0)
        m_ref = factory(pretrained=False)
        saved_qt = ref.model.qt
        ref.model.qt = our_qt                       # the swap: `import encodec_pytorch_b200.quantization as qt`
        try:
            m_our = factory(pretrained=False)
        finally:
            ref.model.qt = saved_qt
    assert type(m_our.quantizer).__module__.startswith("encodec_pytorch_b200")
    assert type(m_ref.quantizer).__module__ == "quantization.vq"
    # random fitted-looking codebooks (pretrained=False leaves kmeans-init zeros): same tables in both models
    g = torch.Generator().manual_seed(7)
    sd = m_ref.state_dict()
    for k in list(sd):
        if k.endswith("_codebook.embed"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * (0.9 ** int(k.split(".")[3]))
            sd[k.replace(".embed", ".embed_avg")] = sd[k].clone()
        elif k.endswith("_codebook.inited"):
            sd[k] = torch.ones_like(sd[k])
        elif k.endswith("_codebook.cluster_size"):
            sd[k] = torch.full_like(sd[k], 10.0)
    m_ref.load_state_dict(sd)
    m_our.load_state_dict(sd)                       # identical keys: quantizer.vq.layers.N._codebook.*
    assert list(m_our.state_dict().keys()) == list(m_ref.state_dict().keys())
    return m_ref.cuda().eval(), m_our.cuda().eval()


@pytest.mark.parametrize("which,bw,channels,seconds", [("24khz", 6.0, 1, 2.0), ("24khz", 24.0, 1, 1.0), ("48khz", 24.0, 2, 2.5)])
def test_encode_decode_through_encodec_model(ref, which, bw, channels, seconds):
    m_ref, m_our = _build_pair(ref, which)
    m_ref.set_target_bandwidth(bw)
    m_our.set_target_bandwidth(bw)
    g = torch.Generator().manual_seed(3)
    wav = (torch.randn(2, channels, int(seconds * m_ref.sample_rate), generator=g) * 0.3).cuda()
    with torch.no_grad():
        f_ref = m_ref.encode(wav)
        f_our = m_our.encode(wav)
    assert len(f_ref) == len(f_our)
    total = diff = 0
    for (c_ref, s_ref), (c_our, s_our) in zip(f_ref, f_our):
        assert c_our.shape == c_ref.shape and c_our.dtype == c_ref.dtype == torch.int64
        assert c_our.stride() == c_ref.stride()                     # [B, K, T] as a transposed view (model.py:166)
        assert (s_ref is None) == (s_our is None)
        if s_ref is not None:
            assert torch.equal(s_ref, s_our)
        # a differing code moves every later stage of that frame: count frames, not codes
        total += c_ref.shape[0] * c_ref.shape[2]
        diff += int((c_ref != c_our).any(dim=1).sum())
    # the reference searches with a cuBLAS fp32 GEMM on the GPU; only fp32 near-ties may move a frame
    assert diff <= max(1, 2e-3 * total), (diff, total)
    with torch.no_grad():
        w_ref = m_ref.decode(f_ref)
        w_our = m_our.decode(f_ref)                                  # same codes in -> same waveform out
    torch.testing.assert_close(w_our, w_ref, rtol=1e-5, atol=1e-5 * float(w_ref.abs().max()))
    # eval forward = decode(encode(x)) cut to the input length (model.py:194-215)
    with torch.no_grad():
        y = m_our(wav)
    assert y.shape == wav.shape and torch.isfinite(y).all()


def test_ecdc_stream_is_byte_identical(ref):
    """compress.py:30-92 (reference, BitPacker loop, CPU ints) against the pack_frame path, fed the same codes."""
    from encodec_pytorch_b200 import compress as our_compress
    for which, bw, channels, seconds in (("24khz", 6.0, 1, 1.5), ("48khz", 12.0, 2, 2.2)):
        m_ref, m_our = _build_pair(ref, which)
        for m in (m_ref, m_our):
            m.set_target_bandwidth(bw)
            m.name = "encodec_" + which                                 # compress.py:46 accepts only the released names
        g = torch.Generator().manual_seed(11)
        wav = (torch.randn(channels, int(seconds * m_ref.sample_rate), generator=g) * 0.3).cuda()
        with torch.no_grad():
            frames = m_our.encode(wav[None])
        # both writers see the same frames (so a near-tie between the two searches cannot blur the byte comparison)
        m_ref.encode = lambda x, frames=frames: frames
        m_our.encode = lambda x, frames=frames: frames
        a, b = io.BytesIO(), io.BytesIO()
        ref.compress.compress_to_file(m_ref, wav, a, use_lm=False)
        our_compress.compress_to_file(m_our, wav, b, use_lm=False)
        assert a.getvalue() == b.getvalue(), which
        if which == "24khz":                                             # one segment: the reference reader applies as is
            del m_our.encode
            w_our, sr = our_compress.decompress(m_our, b.getvalue(), device="cuda")
            w_ref, sr2 = ref.compress.decompress_from_file(m_our, io.BytesIO(a.getvalue()), device="cuda")
            assert sr == sr2 == m_our.sample_rate
            torch.testing.assert_close(w_our.cpu(), w_ref.cpu(), rtol=0, atol=0)
        with pytest.raises(RuntimeError, match="LM entropy coder"):
            our_compress.compress_to_file(m_our, wav, io.BytesIO(), use_lm=True)


def _ddp_worker(rank, world, port, ret):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import install_reference as IR
    import encodec_pytorch_b200.quantization as our_qt
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = IR.import_reference()
            ns.model.qt = our_qt
            torch.manual_seed(0)
            model = ns.model.EncodecModel._get_model([1.5, 3., 6, 12., 24.], 24_000, 1, causal=True, model_norm="weight_norm",
                                                     audio_normalize=False, name="unset").cuda()
            model.train()
            ddp = DDP(model, device_ids=[rank], broadcast_buffers=False, find_unused_parameters=True)   # train_multi_gpu.py:318
            g = torch.Generator().manual_seed(100 + rank)
            x = (torch.randn(2, 1, 24_000, generator=g) * 0.3).cuda()
            out, loss_w, frames = ddp(x)                                        # train_multi_gpu.py:61
            assert out.shape == x.shape and loss_w.shape == (1,) and torch.isfinite(loss_w).all()
            (out.abs().mean() + loss_w.sum()).backward()                        # train_multi_gpu.py:94 (losses reduced to the RVQ term)
            gn = sum(float(p.grad.abs().sum()) for p in model.encoder.parameters() if p.grad is not None)
            assert gn > 0 and all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
            cb = model.quantizer.vq.layers[0]._codebook
            assert float(cb.inited) == 1.0 and float(cb.cluster_size.sum()) > 0   # k-means init + EMA update ran
            # a second step: the steady-state fused path
            out, loss_w, _ = ddp(x)
            (out.abs().mean() + loss_w.sum()).backward()
            torch.cuda.synchronize()
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_training_forward_under_ddp(ref):
    import torch.multiprocessing as mp
    world = min(2, torch.cuda.device_count())
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_ddp_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world
