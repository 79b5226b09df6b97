"""Debug: per-warp event timeline of CTA 0 (first tile, first 4 stages) of the tcgen05 encode."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops, _lib as L
from oracle import cases as CC
torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=128, n_q=32, bins=1024, kmeans_init=False).cuda().eval()
x = CC.latents(64, 128, 750, 1234).cuda()
pk = q.vq._stack_pack()
for _ in range(3): ops.encode(pk, x, 0, 32)
torch.cuda.synchronize()
lib = L.load()
n = 4 * 11 * 16
arr = (C.c_longlong * n)()
lib.rvq_debug_trace.restype = C.c_int
lib.rvq_debug_trace.argtypes = [C.c_void_p, C.c_int]
lib.rvq_debug_trace(arr, n)
names = {0: "start", 9: "E done", 10: "passA", 11: "bar", 12: "resolved", 13: "upd done"}
for s in range(4):
    print(f"--- stage {s}")
    for w in range(11):
        ev = [arr[(s * 11 + w) * 16 + e] for e in range(16)]
        role = "score" if w < 4 else "help " if w < 8 else "tma  " if w == 8 else "mma  "
        print(f"w{w:2d} {role}: " + " ".join(f"{e}:{v}" for e, v in enumerate(ev) if v))
