"""GPU diagnostic: search statistics and timing of the tensor-core encode vs the exact SIMT path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops, _lib as L
from oracle import cases as C

B, T, NQ = int(os.environ.get("B", 64)), 750, int(os.environ.get("NQ", 32))
torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=128, n_q=NQ, bins=1024, kmeans_init=False).cuda().eval()
x = C.latents(B, 128, T, 1234).cuda()
pk = q.vq._stack_pack()
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for nq in (1, 2, 8, NQ):
    t_tc = timed(lambda: ops.encode(pk, x, 0, nq))
    st = ops.search_stats(pk)
    t_ex = timed(lambda: ops.encode(pk, x, 0, nq, flags=L.FLAG_FORCE_EXACT))
    c1 = ops.encode(pk, x, 0, nq)[0]; c2 = ops.encode(pk, x, 0, nq, flags=L.FLAG_FORCE_EXACT)[0]
    print(f"n_q={nq}: tc {t_tc:.3f} ms, exact {t_ex:.3f} ms, stats {st}, code mismatches {(c1 != c2).sum().item()}")
st = ops.search_stats(pk)
w = max(1, st["warps"])
ctas = max(1, st["warps"] // 4)
print("MMA thread per CTA (avg cycles): wait A %d, wait TMA %d, wait acc %d, issue %d, total %d" % tuple(st[k] // ctas for k in ("mma_wait_a", "mma_wait_full", "mma_wait_acc", "mma_issue", "mma_total")))
print("late TMA chunks: %d of %d, avg issue->landed latency %.0f cycles" % (st["tma_late_n"], (B*T+127)//128*NQ*8, st["tma_late_lat_sum"]/max(1,st["tma_late_n"])))
print("per-warp cycles (avg):", {k: st[k] // w for k in st if k.startswith("cyc_")}, "warps", st["warps"])
tiles = (B * T + 127) // 128
print(f"tile-stages per warp ~ {tiles * NQ / (st['warps'] / 4):.1f}; cycles per tile-stage ~ {st['cyc_total'] / w / (tiles * NQ / (st['warps'] / 4)):.0f}")
