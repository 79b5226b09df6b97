import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import _ops as ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import latents, quantizer
dev = torch.device("cuda", 0)
q = quantizer(32, dev, kmeans=True).train()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    with torch.no_grad():
        for i in range(25):
            q(latents(64, 750, 500 + i, dev), 75, 24.0)
q.eval()
x = latents(64, 750, 900, dev)
pk = q.vq._stack_pack()
res = x
with torch.no_grad():
    for i, l in enumerate(q.vq.layers):
        e = l._codebook.embed
        n = e.norm(dim=1)
        cs = l._codebook.cluster_size
        pki = ops.pack([e])
        with ops.search_counters(dev) as counters:
            c, _, _, r = ops.encode(pki, res, 0, 1, want_residual=True)
        st = counters.read()
        rn = res.permute(0, 2, 1).reshape(-1, 128).norm(dim=1)
        print(f"stage {i:2d}: code norms min {n.min():.4f} med {n.median():.4f} max {n.max():.4f}  max|e| {e.abs().max():.3f}  cluster_size min {cs.min():.3f} max {cs.max():.1f} | |r| med {rn.median():.3f} max {rn.max():.3f} | certified {st['certified']} rescored {st['rescored']} fullscan {st['fullscan']}")
        res = r.permute(0, 2, 1)
