"""Encode time at cfg2, cfg4 and 1e6 frames (A/B runs of kernel revisions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sweep
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
res = []
for (b, t, nq) in ((64, 750, 32), (32, 4500, 16), (1333, 750, 32)):
    q = sweep.quantizer(nq, dev).eval()
    n = 1 if b > 1000 else 6
    xs = [sweep.latents(b, t, 1234 + i, dev) for i in range(n)]
    with torch.no_grad():
        ms = sweep.timed(lambda i: q.encode(xs[i % n], 75, None), 10 if b > 1000 else 60)
    res.append(f"[{b},{t}] n_q={nq}: {ms:.4f} ms")
    del xs
print("   ".join(res))
