"""Profiling driver: eval encodes of the trained-like cfg2 stack (k-means init + 25 EMA steps; scripts/sweep.py) for ncu."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sweep
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
q = sweep.quantizer(32, dev, kmeans=True).train()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    with torch.no_grad():
        for i in range(25):
            q(sweep.latents(64, 750, 500 + i, dev), 75, 24.0)
q.eval()
x = sweep.latents(64, 750, 900, dev)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: only the eval encodes below are captured
with torch.no_grad():
    for _ in range(5):
        c = q.encode(x, 75, 24.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", int(c.sum()))
