"""Profiling driver: steady-state training forwards of cfg3 on FITTED tables (k-means init + 25 EMA steps, bench.py's fit).
Run under `ncu --profile-from-start off --metrics gpu__time_duration.sum` for the launch list of the last forwards."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import encodec_pytorch_b200 as E

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
xs = [bench._latents(bench.B, bench.D, bench.T, 1234 + 17 * i).to(dev) for i in range(8)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    torch.manual_seed(0)
    qt = E.ResidualVectorQuantizer(dimension=bench.D, n_q=bench.NQ, bins=bench.BINS, kmeans_init=True, kmeans_iters=10).to(dev).train()
    with torch.no_grad():
        for i in range(26):
            qt(xs[i % 8], bench.FRAME_RATE, bench.BW)
        ms = bench._timed(lambda: qt(xs[3], bench.FRAME_RATE, bench.BW), 20)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(2):
            r = qt(xs[3], bench.FRAME_RATE, bench.BW)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
cs = torch.stack([l._codebook.cluster_size for l in qt.vq.layers])
print(f"training forward on fitted tables: {ms:.3f} ms; largest cluster sizes per stage (EMA): {cs.max(1).values[:8].tolist()}")
