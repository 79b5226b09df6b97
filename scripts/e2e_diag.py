import os, sys, time, json
sys.path.insert(0, "/root/repo")
import torch
import bench
import encodec_pytorch_b200 as E
B, D, T, NQ = 64, 128, 750, 32
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
torch.manual_seed(0)
q = E.ResidualVectorQuantizer(dimension=D, n_q=NQ, bins=1024, kmeans_init=False).to(dev).eval()
nb = 3
t0 = time.perf_counter()
xh = [bench._latents(B, D, T, 99 + i).pin_memory() for i in range(nb)]
ch = [torch.empty((NQ, B, T), dtype=torch.int64).pin_memory() for _ in range(nb)]
print("pin time", time.perf_counter() - t0)
xd = [torch.empty((B, D, T), device=dev) for _ in range(nb)]
s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
ev_in = [torch.cuda.Event() for _ in range(nb)]; ev_cmp = [torch.cuda.Event() for _ in range(nb)]; ev_out = [torch.cuda.Event() for _ in range(nb)]
held = [None] * nb
def e2e_run(n_steps, do_in=True, do_cmp=True, do_out=True):
    for i in range(n_steps):
        k = i % nb
        with torch.cuda.stream(s_in):
            if i >= nb: s_in.wait_event(ev_cmp[k])
            if do_in: xd[k].copy_(xh[k], non_blocking=True)
            ev_in[k].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[k])
            if i >= nb: s_cmp.wait_event(ev_out[k])
            if do_cmp: c = q.encode(xd[k], 75, 24.0)
            ev_cmp[k].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[k])
            if do_out and do_cmp:
                ch[k].copy_(c, non_blocking=True)
            ev_out[k].record(s_out)
        if do_cmp: held[k] = c
def timed(n, **kw):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    a.record(s_in); e2e_run(n, **kw); h1 = time.perf_counter()
    s_out.wait_stream(s_in); s_out.wait_stream(s_cmp); b.record(s_out); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (h1 - h0) / n * 1e3
with torch.no_grad():
    e2e_run(6); torch.cuda.synchronize()
    for rep in range(4):
        print("e2e rep", rep, "gpu ms/step %.3f  host enqueue ms/step %.3f" % timed(100))
    print("in only   %.3f host %.3f" % timed(100, do_cmp=False, do_out=False))
    print("cmp only  %.3f host %.3f" % timed(100, do_in=False, do_out=False))
    print("cmp+out   %.3f host %.3f" % timed(100, do_in=False))
    print("e2e again %.3f host %.3f" % timed(100))
