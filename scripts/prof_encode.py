"""Profiling driver: a few fused encodes of BASELINE cfg2 (latents [64,128,750], n_q=32) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E
from oracle import cases as C

B, T, NQ = int(os.environ.get("B", 64)), int(os.environ.get("T", 750)), int(os.environ.get("NQ", 32))
torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=128, n_q=NQ, bins=1024, kmeans_init=False).cuda().eval()
x = C.latents(B, 128, T, 1234).cuda()
with torch.no_grad():
    for _ in range(int(os.environ.get("REPS", 4))):
        c = q.encode(x, 75, None)
torch.cuda.synchronize()
print("ok", int(c.sum()))
