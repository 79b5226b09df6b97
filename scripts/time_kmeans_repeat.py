"""k-means init step of cfg3 (8 stages x 50 Lloyd iterations on 48 000 frames), repeated with different seeds: time and what the
assignment searches did (the step is bimodal: ~145 ms or > 1 s)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sweep
from encodec_pytorch_b200 import _ops as ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
x = sweep.latents(64, 750, 1234, dev)
for seed in range(int(os.environ.get("SEEDS", 6))):
    torch.manual_seed(seed)
    qk = sweep.quantizer(8, dev, kmeans=True).train()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with ops.search_counters(dev) as c, torch.no_grad():
            a.record()
            qk(x, 75, 6.0)
            e.record()
            torch.cuda.synchronize()
    print(seed, f"{a.elapsed_time(e):.1f} ms", c.read(), flush=True)
