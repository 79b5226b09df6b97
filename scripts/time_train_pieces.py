"""Device time of the pieces of a steady-state training forward on fitted tables, timed inside the sequence they run in."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops, _lib as L

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
xs = [bench._latents(bench.B, bench.D, bench.T, 1234 + 17 * i).to(dev) for i in range(8)]
def ev(): return torch.cuda.Event(enable_timing=True)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    torch.manual_seed(0)
    fit = os.environ.get("FIT", "1") == "1"
    qt = E.ResidualVectorQuantizer(dimension=bench.D, n_q=bench.NQ, bins=bench.BINS, kmeans_init=fit, kmeans_iters=10).to(dev).train()
    with torch.no_grad():
        for i in range(26 if fit else 2):
            qt(xs[i % 8], bench.FRAME_RATE, bench.BW)
        cbs = [l._codebook for l in qt.vq.layers]
        x = xs[3]
        acc = {}
        for it in range(12):
            e = [ev() for _ in range(6)]
            e[0].record()
            pk = ops.pack([cb.embed for cb in cbs])
            e[1].record()
            codes, _, sqerr, res = ops.encode(pk, x, 0, 32, want_sqerr=True, want_residual=True, flags=L.FLAG_STE)
            e[2].record()
            quant = x.permute(0, 2, 1) - res
            e[3].record()
            flat, counts, esum = ops.ema_stats(pk, x, codes, 0, L.FLAG_STE)
            e[4].record()
            ops.ema_apply([cb.cluster_size for cb in cbs], [cb.embed_avg for cb in cbs], [cb.embed for cb in cbs], counts, esum, 0.99, 1e-5)
            e[5].record()
            torch.cuda.synchronize()
            if it >= 2:
                for k, name in enumerate(("pack", "search", "quant", "stats", "apply")):
                    acc[name] = acc.get(name, 0.) + e[k].elapsed_time(e[k + 1]) / 10
        print({k: round(v, 4) for k, v in acc.items()}, "sum", round(sum(acc.values()), 3))
        ms = bench._timed(lambda: qt(x, bench.FRAME_RATE, bench.BW), 20)
        print("module forward", round(ms, 3))
        # the same pieces back to back without events in between
        def seq():
            pk = ops.pack([cb.embed for cb in cbs])
            codes, _, sqerr, res = ops.encode(pk, x, 0, 32, want_sqerr=True, want_residual=True, flags=L.FLAG_STE)
            quant = x.permute(0, 2, 1) - res
            flat, counts, esum = ops.ema_stats(pk, x, codes, 0, L.FLAG_STE)
            ops.ema_apply([cb.cluster_size for cb in cbs], [cb.embed_avg for cb in cbs], [cb.embed for cb in cbs], counts, esum, 0.99, 1e-5)
        print("pieces back to back", round(bench._timed(seq, 20), 3))
