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
print({k: st[k] for k in ("searched", "certified", "rescored", "fullscan")})
