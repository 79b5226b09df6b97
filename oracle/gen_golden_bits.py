"""Generate tests/golden/bitpack_*.npz by running the UNMODIFIED reference's binary.py on CPU.

TEST INFRASTRUCTURE ONLY (build container only; /root/reference does not travel):

    python -m oracle.gen_golden_bits

For each case: seeded random codes ``frame [1, K, T]`` (values < 2**bits) pushed through ``binary.BitPacker`` in the loop of
``compress.compress_to_file`` (compress.py:70-92), the resulting bytes, and the values ``binary.BitUnpacker`` pulls back.
"""
from __future__ import annotations

import io
import os
import sys

import numpy as np

REF = os.environ.get("RVQ_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, K, T, bits, seed)
CASES = [
    ("cfg1_k8_t750_b10", 8, 750, 10, 1),
    ("k32_t75_b10", 32, 75, 10, 2),
    ("k16_t150_b10", 16, 150, 10, 3),
    ("odd_k3_t37_b7", 3, 37, 7, 4),
    ("one_k1_t1_b10", 1, 1, 10, 5),
    ("k2_t5_b1", 2, 5, 1, 6),
    ("k5_t33_b16", 5, 33, 16, 7),
    ("k8_t1001_b11", 8, 1001, 11, 8),
]


def codes_for(k: int, t: int, bits: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 2 ** bits, size=(1, k, t), dtype=np.int64)


def main() -> None:
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    sys.path.insert(0, REF)
    import binary  # the reference's binary.py
    os.makedirs(OUT, exist_ok=True)
    for name, k, t, bits, seed in CASES:
        frame = codes_for(k, t, bits, seed)
        fo = io.BytesIO()
        packer = binary.BitPacker(bits, fo)
        for ti in range(t):                                   # compress.py:82-90
            for value in frame[0, :, ti].tolist():
                packer.push(value)
        packer.flush()
        data = fo.getvalue()
        fo.seek(0)
        unpacker = binary.BitUnpacker(bits, fo)
        pulled = [unpacker.pull() for _ in range(k * t)]
        assert None not in pulled
        np.savez_compressed(os.path.join(OUT, f"bitpack_{name}.npz"), k=k, t=t, bits=bits, seed=seed,
                            data=np.frombuffer(data, dtype=np.uint8), pulled=np.asarray(pulled, dtype=np.int64))
        print(name, len(data), "bytes")


if __name__ == "__main__":
    main()
