"""Decode / eval-forward time at cfg2 and cfg4 (chain kernels: L2-gather bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sweep
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
for name, b, t, nq, fr in (("cfg2", 64, 750, 32, 75), ("cfg4", 32, 4500, 16, 150)):
    q = sweep.quantizer(nq, dev).eval()
    xs = [sweep.latents(b, t, 10 + i, dev) for i in range(4)]
    with torch.no_grad():
        codes = [q.encode(x, fr, 24.0) for x in xs]
        ms_d = sweep.timed(lambda i: q.decode(codes[i % 4]), 50)
        q.contiguous_outputs = True
        ms_dc = sweep.timed(lambda i: q.decode(codes[i % 4]), 50)
        q.contiguous_outputs = False
        ms_f = sweep.timed(lambda i: q(xs[i % 4], fr, 24.0), 30)
        ms_e = sweep.timed(lambda i: q.encode(xs[i % 4], fr, 24.0), 30)
    gath = b * t * nq * 512 / 1e9
    print(f"{name}: decode {ms_d:.4f} ms ({gath / ms_d * 1e3:.0f} GB/s of L2 gathers), decode [B,D,T] {ms_dc:.4f} ms, eval forward {ms_f:.4f} ms (encode {ms_e:.4f})", flush=True)
