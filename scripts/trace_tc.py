"""Timeline of CTA 0 of the tcgen05 search (library built with RVQ_NVCC_DEFS=RVQ_TC_TRACE): per slot and step the cycle
stamps of the hand-offs MMA -> score -> update -> MMA."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops, _lib as L
from oracle import cases as Cs

torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=128, n_q=32, bins=1024, kmeans_init=False).cuda().eval()
x = Cs.latents(64, 128, 750, 1234).cuda()
pk = q.vq._stack_pack()
for _ in range(3): ops.encode(pk, x, 0, 32)
torch.cuda.synchronize()
lib = L.load()
STEPS, EV = 6, 16
N0 = int(os.environ.get('TRACE_N0', 2))   # must match -DRVQ_TRACE_N0 of the build
arr = (C.c_longlong * (2 * STEPS * EV + 128 + 128))()
lib.rvq_debug_trace.restype = C.c_int
lib.rvq_debug_trace.argtypes = [C.c_void_p, C.c_int]
lib.rvq_debug_trace(arr, len(arr))
names = ["mma:A seen", "mma:chunk0 issued", "mma:all issued", "score:first acc", "score:chunks done", "score:cand arrive",
         "upd:cand seen", "upd:pass done", "upd:A arrive"]
for n in range(STEPS):
    for X in range(2):
        ev = [arr[(X * STEPS + n) * EV + e] for e in range(9)]
        print(f"step {n + N0} slot {X}: " + "  ".join(f"{nm}={v}" for nm, v in zip(names, ev)))
base = 2 * STEPS * EV
print("MMA thread, slot 0 step N0+2, per chunk: acc free | third0 landed | third1 | third2 | all issued")
for c in range(8):
    print(f"  chunk {c}: " + "  ".join(str(arr[base + 8 * c + e]) for e in range(5)))
print("score warp 0, slot 0 step N0+2, per chunk: acc_full seen | accumulator released | minima done")
for c in range(8):
    print(f"  chunk {c}: " + "  ".join(str(arr[base + 64 + 3 * c + e]) for e in range(3)))
for X in range(2):
    print(f"update warps, slot {X} step N0+2: start | resolve done | barrier passed | rows requested | frame A done | frame B done | operand stored")
    for u in range(8):
        print(f"  warp {u}: " + "  ".join(str(arr[base + 128 + 64 * X + 8 * u + e]) for e in range(8)))
