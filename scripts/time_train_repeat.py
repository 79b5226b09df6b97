import os, sys, warnings, time
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
    for i in range(26):
        qt(xs[i % 8], bench.FRAME_RATE, bench.BW)
    for rep in range(8):
        seg0 = torch.cuda.memory_stats()["num_device_alloc"]
        if rep % 2 == 0:
            ms = bench._timed(lambda: qt(xs[3], bench.FRAME_RATE, bench.BW), 20)
        else:
            with torch.no_grad():
                ms = bench._timed(lambda: qt(xs[3], bench.FRAME_RATE, bench.BW), 20)
        print(rep, "grad" if rep % 2 == 0 else "no_grad", round(ms, 3), "cudaMallocs", torch.cuda.memory_stats()["num_device_alloc"] - seg0, flush=True)
        from encodec_pytorch_b200 import _ops as ops, _lib as L
        with torch.no_grad():
            pk = qt.vq._stack_pack()
            with ops.search_counters(dev) as c:
                codes, _, sqerr, res = ops.encode(pk, xs[3], 0, 32, want_sqerr=True, want_residual=True, flags=L.FLAG_STE)
            st = c.read()
            t_s = bench._timed(lambda: ops.encode(pk, xs[3], 0, 32, want_sqerr=True, want_residual=True, flags=L.FLAG_STE), 10)
            t_e = bench._timed(lambda: ops.ema_stats(pk, xs[3], codes, 0, L.FLAG_STE), 10)
            nrm = torch.stack([l._codebook.embed.norm(dim=1) for l in qt.vq.layers])
            srt = nrm.sort(dim=1).values
            print("   norm quantiles stage 1/16/31 (min, 10%, 50%, 90%, max):", [[round(float(srt[i, j]), 3) for j in (0, 102, 512, 921, 1023)] for i in (1, 16, 31)])
            big = torch.stack([torch.bincount(codes[i].reshape(-1), minlength=1024).max() for i in range(32)])
            print("   search", round(t_s, 3), "stats", round(t_e, 3), st, "largest cluster per stage (first 6)", big[:6].tolist(), "max", int(big.max()), flush=True)
