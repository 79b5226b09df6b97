"""The bit-packing oracle against the byte streams produced by the reference's binary.py (CPU)."""
import glob
import os

import numpy as np
import pytest

from oracle import binary_oracle as BO
from oracle.gen_golden_bits import CASES, codes_for

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_every_case_has_a_fixture():
    have = {os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "bitpack_*.npz"))}
    assert have == {f"bitpack_{c[0]}.npz" for c in CASES}


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_oracle_matches_reference_bytes(case):
    name, k, t, bits, seed = case
    g = np.load(os.path.join(GOLD, f"bitpack_{name}.npz"))
    frame = codes_for(k, t, bits, seed)
    ref = g["data"]
    assert ref.size == BO.packed_nbytes(k * t, bits)
    values = frame[0].T.reshape(-1)
    assert np.array_equal(np.frombuffer(BO.pack_values_loop(values.tolist(), bits), dtype=np.uint8), ref)
    assert np.array_equal(BO.pack_values(values, bits), ref)
    assert np.array_equal(BO.pack_frame(frame, bits)[0], ref)
    assert np.array_equal(g["pulled"], values)
    assert BO.unpack_values_loop(ref.tobytes(), bits, k * t) == values.tolist()
    assert np.array_equal(BO.unpack_frame(ref[None], k, t, bits), frame)


def test_empty_and_batched():
    assert BO.pack_frame(np.zeros((2, 4, 0), np.int64), 10).shape == (2, 0)
    f = np.random.default_rng(0).integers(0, 1024, size=(3, 8, 21), dtype=np.int64)
    p = BO.pack_frame(f, 10)
    assert p.shape == (3, BO.packed_nbytes(8 * 21, 10))
    assert np.array_equal(BO.unpack_frame(p, 8, 21, 10), f)
