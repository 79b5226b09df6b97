"""GPU diagnostic of the two-tile tcgen05 search: timing, search counters and agreement with the exact fp32 path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops, _lib as L
from oracle import cases as C

B, T, NQ = int(os.environ.get("B", 64)), int(os.environ.get("T", 750)), int(os.environ.get("NQ", 32))
torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=128, n_q=NQ, bins=1024, kmeans_init=False).cuda().eval()
x = C.latents(B, 128, T, 1234).cuda()
pk = q.vq._stack_pack()
def timed(fn, n=10):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
import bench as _bench
for nq in sorted({1, 8, NQ}):
    sampler = _bench.ClockSampler(_bench._physical_gpu_index(0), period=0.002)
    sampler.start()
    t_tc = timed(lambda: ops.encode(pk, x, 0, nq), n=10 if nq < NQ else 100)
    clk = sampler.stop()
    with ops.search_counters(x.device) as counters:
        c1 = ops.encode(pk, x, 0, nq)[0]
    st = counters.read()
    c2 = ops.encode(pk, x, 0, nq, flags=L.FLAG_FORCE_EXACT)[0]
    N = B * T
    tiles = (N + 127) // 128
    ctas = min(148, tiles)
    ts = tiles * nq                      # tile-stages in the launch
    print(f"n_q={nq}: tc {t_tc:.3f} ms ({t_tc*1e-3*1.965e9/ (-(-tiles//ctas)*nq):.0f} cyc per tile-stage of the longest CTA), "
          f"certified {st['certified']}/{st['searched']}, rescored {st['rescored']}, fullscan {st['fullscan']}, mismatches vs exact {(c1 != c2).sum().item()}, sm clock {clk['sm_mhz']} MHz {clk['reasons']}")
# the training variant of the kernel (straight-through arithmetic, loss numerators): which option costs what
stats = ops.ema_stats_buffer(NQ, pk.K, pk.D, x.device)
for name, kw in (("ste+sqerr", dict(want_sqerr=True, flags=L.FLAG_STE)), ("ste only", dict(flags=L.FLAG_STE)), ("sqerr only", dict(want_sqerr=True)),
                 ("residual only", dict(want_residual=True)), ("stats only", dict(ema_stats_out=stats)),
                 ("full training call", dict(want_sqerr=True, want_residual=True, flags=L.FLAG_STE, ema_stats_out=stats))):
    t = timed(lambda: ops.encode(pk, x, 0, NQ, **kw), n=50)
    print(f"train variant, {name}: {t:.3f} ms")
