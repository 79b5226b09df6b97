"""GPU parity of the code bit-packing (rvq_bitpack / rvq_bitunpack through encodec_pytorch_b200.binary) against the byte
streams of the reference's binary.py (golden fixtures) and against the CPU oracle.  Bit-exact: integer/byte work."""
import os

import numpy as np
import pytest
import torch

from oracle import binary_oracle as BO
from oracle.gen_golden_bits import CASES, codes_for

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_matches_reference_byte_streams(case):
    from encodec_pytorch_b200 import binary as B
    name, k, t, bits, seed = case
    g = np.load(os.path.join(GOLD, f"bitpack_{name}.npz"))
    frame = torch.from_numpy(codes_for(k, t, bits, seed)).cuda()
    packed = B.pack_frame(frame, bits)
    assert packed.shape == (1, B.packed_nbytes(k, t, bits))
    assert np.array_equal(packed[0].cpu().numpy(), g["data"])
    back = B.unpack_frame(torch.from_numpy(g["data"].copy())[None].cuda(), k, t, bits)
    assert torch.equal(back, frame)


@pytest.mark.parametrize("bits", [1, 3, 7, 8, 10, 11, 13, 16])
@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 8, 75), (2, 32, 129), (5, 17, 1000), (2, 64, 257)])
def test_random_frames_against_the_oracle(bits, shape):
    from encodec_pytorch_b200 import binary as B
    b, k, t = shape
    rng = np.random.default_rng(1000 * bits + b + k + t)
    f = rng.integers(0, 2 ** bits, size=(b, k, t), dtype=np.int64)
    want = BO.pack_frame(f, bits)
    fd = torch.from_numpy(f).cuda()
    got = B.pack_frame(fd, bits)
    assert np.array_equal(got.cpu().numpy(), want)
    # the search's own layout: [K, B, T] storage viewed as [B, K, T] (model.py:165-166)
    kbt = fd.permute(1, 0, 2).contiguous()
    assert torch.equal(B.pack_frame(kbt.permute(1, 0, 2), bits), got)
    assert torch.equal(B.unpack_frame(got, k, t, bits), fd)
    # longer rows than needed (a stream embedded in a larger buffer) unpack the same
    padded = torch.cat([got, torch.full((b, 5), 255, dtype=torch.uint8, device="cuda")], 1)
    assert torch.equal(B.unpack_frame(padded, k, t, bits), fd)


def test_edges_and_errors():
    from encodec_pytorch_b200 import binary as B
    assert B.pack_frame(torch.zeros((2, 4, 0), dtype=torch.int64, device="cuda"), 10).shape == (2, 0)
    assert B.pack_frame(torch.zeros((0, 4, 9), dtype=torch.int64, device="cuda"), 10).shape == (0, 45)
    assert B.unpack_frame(torch.zeros((2, 0), dtype=torch.uint8, device="cuda"), 4, 0, 10).shape == (2, 4, 0)
    # values are masked to `bits` bits (the reference adds them unmasked, which corrupts the neighbours: binary.py:72)
    f = torch.tensor([[[1023 + 1024, 5]]], dtype=torch.int64, device="cuda")
    assert torch.equal(B.unpack_frame(B.pack_frame(f, 10), 1, 2, 10), torch.tensor([[[1023, 5]]], device="cuda"))
    with pytest.raises(RuntimeError):
        B.pack_frame(torch.zeros((1, 2, 3), dtype=torch.int64), 10)              # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        B.pack_frame(torch.zeros((1, 2, 3), dtype=torch.int64, device="cuda"), 17)
    with pytest.raises(RuntimeError):
        B.unpack_frame(torch.zeros((1, 3), dtype=torch.uint8, device="cuda"), 2, 3, 10)   # 8 bytes needed


def test_full_size_round_trip_from_the_search():
    """cfg2 / cfg4 sizes: pack the codes of a real encode, unpack, compare; spot-check streams against the oracle."""
    import encodec_pytorch_b200 as E
    from encodec_pytorch_b200 import binary as B
    torch.manual_seed(0)
    for (b, t, n_q) in ((64, 750, 32), (32, 4500, 16)):
        q = E.ResidualVectorQuantizer(dimension=128, n_q=n_q, bins=1024, kmeans_init=False).cuda().eval()
        x = torch.randn(b, 128, t, device="cuda")
        with torch.no_grad():
            codes = q.encode(x, 75, None)                    # [n_q, B, T]
        frame = codes.transpose(0, 1)                        # [B, K, T] as model.py:166 hands it on
        packed = B.pack_frame(frame, 10)
        assert packed.shape == (b, n_q * t * 10 // 8)
        assert torch.equal(B.unpack_frame(packed, n_q, t, 10), frame)
        for i in (0, b - 1):
            assert np.array_equal(packed[i].cpu().numpy(), BO.pack_values(frame[i].T.reshape(-1).cpu().numpy(), 10))
