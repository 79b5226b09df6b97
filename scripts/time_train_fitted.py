"""Where the steady-state training forward on FITTED tables spends its time (bench.py's fit)."""
import os, sys, warnings, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops, _lib as L

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
xs = [bench._latents(bench.B, bench.D, bench.T, 1234 + 17 * i).to(dev) for i in range(8)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    torch.manual_seed(0)
    qt = E.ResidualVectorQuantizer(dimension=bench.D, n_q=bench.NQ, bins=bench.BINS, kmeans_init=True, kmeans_iters=10).to(dev).train()
    for i in range(26):
        qt(xs[i % 8], bench.FRAME_RATE, bench.BW)
    ms_grad = bench._timed(lambda: qt(xs[3], bench.FRAME_RATE, bench.BW), 20)
    with torch.no_grad():
        ms_nograd = bench._timed(lambda: qt(xs[3], bench.FRAME_RATE, bench.BW), 20)
    # host time of one call (enqueue only)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): qt(xs[3], bench.FRAME_RATE, bench.BW)
    host = (time.perf_counter() - t0) / 20 * 1e3
    torch.cuda.synchronize()
    pk = qt.vq._stack_pack()
    with ops.search_counters(dev) as c:
        ops.encode(pk, xs[3], 0, 32, want_sqerr=True, want_residual=True, flags=L.FLAG_STE)
    st = c.read()
    t_train = bench._timed(lambda: ops.encode(pk, xs[3], 0, 32, want_sqerr=True, want_residual=True, flags=L.FLAG_STE), 20)
    t_eval = bench._timed(lambda: ops.encode(pk, xs[3], 0, 32), 20)
    t_pack = bench._timed(lambda: ops.pack([l._codebook.embed for l in qt.vq.layers]), 20)
print(f"forward with grad mode on {ms_grad:.3f} ms, under no_grad {ms_nograd:.3f} ms, host enqueue {host:.3f} ms")
print(f"search kernel on these tables: train variant {t_train:.3f} ms, eval variant {t_eval:.3f} ms, certified {st['certified']/st['searched']:.4f}; pack {t_pack:.3f} ms")
import numpy as np
print("counters", st)
buf = pk.buf.cpu().numpy()
stride, off_meta = 1359104, 1347584
for s in range(32):
    m = buf[256 + s * stride + off_meta: 256 + s * stride + off_meta + 44]
    f = m.view(np.float32); i = m.view(np.int32)
    if s % 4 == 0 or s < 3:
        print(f"stage {s}: coef {f[0]:.3e} abs {f[1]:.2e} xlimit {f[2]:.3g} cref {f[3]:.3g} cmin {f[4]:.3g} nout {i[5]} cmax {f[6]:.3g} mdr {f[7]:.3g} percode {i[8]} abs_pc {f[9]:.2e} g16max {f[10]:.3e}")
# time per stage prefix
for nq in (1, 2, 4, 8, 16, 32):
    t = bench._timed(lambda: ops.encode(pk, xs[3], 0, nq), 20)
    with ops.search_counters(dev) as c:
        ops.encode(pk, xs[3], 0, nq)
    print(nq, f"{t:.3f} ms", c.read())
