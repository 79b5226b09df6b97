"""cfg3: RVQ training forward (fused search, commitment losses, quantized sum, EMA statistics + all-reduce + apply, pack
refresh) frame-sharded over the ranks of one node, 64 batch items of 750 frames per rank (weak scaling), n_q = 32.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/train_scaling.py [--out f.json]

Reports (rank 0, one JSON line): step time (CUDA events, max over ranks) with the cross-rank buffer sync on
(`distrib.sync_buffers(True)`: one 16.1 MiB NCCL all-reduce of the packed EMA statistics per step) and off (the
reference's literal behaviour: no collective), the stand-alone time of that all-reduce, and forward+backward."""
import argparse, json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import distrib

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--out", default=None)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
B, D, T, NQ = 64, 128, 750, 32


def lat(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, D, T, generator=g).to(dev)


xs = [lat(1000 + 31 * rank + i) for i in range(6)]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, n):
    for i in range(5):
        fn(i)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out = {"world": world, "frames_per_rank_per_step": B * T, "n_q": NQ}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for name, sync, fitted in (("random_init_sync", True, False), ("random_init_nosync", False, False), ("fitted_sync", True, True)):
        distrib.sync_buffers(sync)
        torch.manual_seed(0)
        q = E.ResidualVectorQuantizer(dimension=D, n_q=NQ, bins=1024, kmeans_init=fitted, kmeans_iters=10).to(dev).train()
        with torch.no_grad():
            for i in range(26 if fitted else 3):
                q(xs[i % 6], 75, 24.0)
            ms = timed(lambda i: q(xs[i % 6], 75, 24.0), args.steps)
        out[name] = {"step_ms": ms, "frames_per_s": world * B * T / (ms * 1e-3)}
        if name == "random_init_sync":
            xg = [x.clone().requires_grad_(True) for x in xs[:2]]

            def fb(i):
                r = q(xg[i % 2], 75, 24.0)
                (r.quantized.sum() + r.penalty).backward()
            out["random_init_sync"]["fwd_bwd_ms"] = timed(fb, max(10, args.steps // 2))
    if world > 1:
        buf = torch.zeros(NQ * 1024 * (D + 1), device=dev)
        out["allreduce_16MiB_ms"] = timed(lambda i: dist.all_reduce(buf), 50)
        # the buffers really are in sync after the synced run
        cs = q.vq.layers[0]._codebook.embed.clone()
        ref = cs.clone()
        dist.broadcast(ref, src=0)
        out["embed_equal_across_ranks"] = bool(torch.equal(cs, ref))
if rank == 0:
    line = json.dumps(out)
    print(line)
    if args.out:
        open(args.out, "w").write(line + "\n")
if world > 1:
    dist.destroy_process_group()
