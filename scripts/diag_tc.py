"""GPU diagnostic of the two-tile tcgen05 search: timing vs the exact path and, when the library was built with
RVQ_NVCC_DEFS=RVQ_TC_TIMERS, the per-role cycle counters (averages per CTA / per tile-stage)."""
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
    st = ops.search_stats(pk)
    c1 = ops.encode(pk, x, 0, nq)[0]; c2 = ops.encode(pk, x, 0, nq, flags=L.FLAG_FORCE_EXACT)[0]
    N = B * T
    tiles = (N + 127) // 128
    ctas = min(148, tiles)
    ts = tiles * nq                      # tile-stages in the launch
    print(f"n_q={nq}: tc {t_tc:.3f} ms ({t_tc*1e-3*1.965e9/ (-(-tiles//ctas)*nq):.0f} cyc per tile-stage of the longest CTA), "
          f"certified {st['certified']}/{st['searched']}, rescored {st['rescored']}, fullscan {st['fullscan']}, mismatches vs exact {(c1 != c2).sum().item()}, sm clock {clk['sm_mhz']} MHz {clk['reasons']}")
    if st["warps"]:
        w = st["warps"]                  # score warps counted = 4 per CTA
        per = lambda k, div: st[k] / div
        print(f"   score warp / tile-stage: wait acc {per('cyc_wait', w)*ctas/ts:.0f}  ld+min {per('cyc_scores', w)*ctas/ts:.0f}  winner {per('cyc_winner', w)*ctas/ts:.0f}  (total per warp {per('cyc_total', w):.0f} cyc)")
        uw = 2 * w                       # update warps = 8 per CTA
        print(f"   update warp / tile-stage: wait cand {per('cyc_resolve', uw)*ctas/ts:.0f}  update {per('cyc_update', uw)*ctas/ts:.0f}  operand->tmem {per('cyc_pairbar', uw)*ctas/ts:.0f} + st wait {per('tma_late_lat_sum', uw)*ctas/ts:.0f}  tile loads (per CTA) {per('cyc_load', uw):.0f}")
        print(f"      update split: setup {per('cand2', uw)*ctas/ts:.0f}  own frames {per('cand3_4', uw)*ctas/ts:.0f}  lists {per('cand5_8', uw)*ctas/ts:.0f}  wide {per('cand9plus', uw)*ctas/ts:.0f}  barrier {per('tma_late_n', uw)*ctas/ts:.0f}")
        print(f"   mma thread / tile-stage: wait A {st['mma_wait_a']/ts:.0f}  wait TMA {st['mma_wait_full']/ts:.0f}  wait acc {st['mma_wait_acc']/ts:.0f}  issue {st['mma_issue']/ts:.0f}  (total per CTA {st['mma_total']/ctas:.0f} cyc)")
# the training variant of the kernel (straight-through arithmetic, loss numerators): which option costs what
for name, kw in (("ste+sqerr", dict(want_sqerr=True, flags=L.FLAG_STE)), ("ste only", dict(flags=L.FLAG_STE)), ("sqerr only", dict(want_sqerr=True))):
    t = timed(lambda: ops.encode(pk, x, 0, NQ, **kw), n=50)
    print(f"train variant, {name}: {t:.3f} ms")
