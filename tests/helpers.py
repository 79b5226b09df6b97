"""Shared helpers of the parity tests (oracle = checker only; see oracle/rvq_oracle.py)."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import cases as C
from oracle import rvq_oracle as O


def load_golden(golden_dir, kind, case):
    return np.load(os.path.join(golden_dir, f"{kind}_{case.name}.npz"))


def build_module(case, kmeans_init=False, device="cuda", **kw):
    """Our drop-in module with the tables the reference constructor draws under the same seed
    (same RNG order: one kaiming_uniform_ per stage, stage 0 first)."""
    import encodec_pytorch_b200 as E
    torch.manual_seed(case.cb_seed)
    q = E.ResidualVectorQuantizer(dimension=case.d, n_q=case.n_q, bins=case.k, kmeans_init=kmeans_init, **kw)
    return q.to(device)


def module_states(q):
    return O.states_from_module(q)


def check_summary(t, g, prefix, stride=97, rtol=0.0, atol=0.0):
    c = t.detach().cpu().contiguous()
    assert list(c.shape) == list(g[prefix + "_shape"]), (list(c.shape), list(g[prefix + "_shape"]))
    if C.sha(c) == str(g[prefix + "_sha"]):
        return "bitexact"
    assert rtol > 0 or atol > 0, f"{prefix}: expected bit-exact output"
    sub = c.reshape(-1)[::stride].numpy()
    np.testing.assert_allclose(sub, g[prefix + "_sub"], rtol=rtol, atol=atol)
    ref_abs = float(g[prefix + "_abs64"])
    assert abs(c.double().abs().sum().item() - ref_abs) <= max(rtol, 1e-6) * ref_abs + atol * c.numel()
    return "close"


def assert_codes_match(states, x_cpu, got, want, rel_gap=1e-6, max_near_frac=2e-3):
    """``got`` vs the reference's ``want`` ([n_q, B, T]): equal, or differing only at documented
    fp32 near-ties (fp64 distance gap < rel_gap relative), judged stage-wise on ``got``'s own
    residual chain.  Returns the teacher-forced statistics."""
    got = got.detach().cpu().to(torch.long)
    want = torch.as_tensor(np.asarray(want).astype(np.int64))
    assert got.shape == want.shape
    if torch.equal(got, want):
        return {"pairs": got.numel(), "mismatch": 0, "near_tie": 0, "bad": 0}
    st = O.compare_codes_teacher_forced(states, x_cpu, got, rel_gap)
    assert st["bad"] == 0, st
    assert st["near_tie"] <= max(2, max_near_frac * st["pairs"]), st
    return st
