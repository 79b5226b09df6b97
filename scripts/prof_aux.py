"""Profiling driver for the HBM/L2-bound kernels at cfg2: decode gather, EMA statistics, EMA apply, expiry, bit-packing."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import binary as BN

torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=128, n_q=32, bins=1024, kmeans_init=False).cuda().eval()
g = torch.Generator().manual_seed(7)
x = torch.randn(64, 128, 750, generator=g).cuda()
with torch.no_grad():
    codes = q.encode(x, 75, None)
    for _ in range(2):
        y = q.decode(codes)
        p = BN.pack_frame(codes.transpose(0, 1), 10)
        c2 = BN.unpack_frame(p, 32, 750, 10)
    q.train()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(2):
            r = q(x, 75, 24.0)
torch.cuda.synchronize()
print("ok", float(y.abs().sum()), bool((c2 == codes.transpose(0, 1)).all()))
