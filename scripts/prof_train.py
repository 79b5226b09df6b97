"""Profiling driver: training forwards (search + quantized + losses + EMA update + expiry) of BASELINE cfg3 for the ncu launch list."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E

torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=128, n_q=32, bins=1024, kmeans_init=False).cuda().train()
g = torch.Generator().manual_seed(7)
x = torch.randn(64, 128, 750, generator=g).cuda()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    with torch.no_grad():
        for _ in range(int(os.environ.get("REPS", 3))):
            r = q(x, 75, 24.0)
torch.cuda.synchronize()
print("ok", float(r.penalty))
