"""Encode time and certified share of cfg2 under the three settings of rvq_pack_bound_mode, on bench.py's fitted stack
(k-means init with 10 iterations + 25 EMA forwards on the bench latents) and on the random-init stack."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
xs = [bench._latents(bench.B, bench.D, bench.T, 1234 + 17 * i).to(dev) for i in range(8)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    torch.manual_seed(0)
    qt = E.ResidualVectorQuantizer(dimension=bench.D, n_q=bench.NQ, bins=bench.BINS, kmeans_init=True, kmeans_iters=10).to(dev).train()
    for i in range(26):
        qt(xs[i % 8], bench.FRAME_RATE, bench.BW)
    qt.eval()
torch.manual_seed(0)
qr = E.ResidualVectorQuantizer(dimension=bench.D, n_q=bench.NQ, bins=bench.BINS, kmeans_init=False).to(dev).eval()
for name, q in (("fitted", qt), ("random-init", qr)):
    for mode in (2, 1, 0):
        with ops.pack_bound_mode(mode):
            q.vq.invalidate()
            with ops.search_counters(dev) as c, torch.no_grad():
                q.encode(xs[0], bench.FRAME_RATE, bench.BW)
            st = c.read()
            with torch.no_grad():
                ms = bench._timed(lambda: q.encode(xs[1], bench.FRAME_RATE, bench.BW), 50)
        print(f"{name:12s} mode {mode}: {ms:.3f} ms  certified {st['certified'] / st['searched']:.4f}  rescored {st['rescored']}  fullscan {st['fullscan']}", flush=True)
