"""Encode time of the trained-like cfg2 stack (sweep.trained_like) and of the random-init one, for A/B runs."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sweep
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
out = {}
sweep.trained_like("t", 64, 750, 32, 75, 24.0, dev, 30, out)
q = sweep.quantizer(32, dev).eval()
xs = [sweep.latents(64, 750, 1234 + i, dev) for i in range(6)]
with torch.no_grad():
    ms = sweep.timed(lambda i: q.encode(xs[i % 6], 75, 24.0), 50)
print(f"trained-like {out['t']['encode_ms']:.3f} ms (certified {out['t']['certified_share']:.3f})   random-init {ms:.3f} ms")
