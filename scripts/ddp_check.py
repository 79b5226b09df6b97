"""Multi-GPU equivalence of the training path on real hardware: P ranks, each on its shard of the batch,
with ONE NCCL all-reduce of the packed EMA statistics per step (and the rank-0 broadcast after the k-means init), must end
up with the codebooks one rank gets on the concatenated batch.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/ddp_check.py

Every rank runs the sharded job; rank 0 then repeats the steps alone on the full batch (process group left untouched but
bypassed by world-size-1 helpers is not possible, so the single-rank reference runs FIRST, before init_process_group)."""
import os, sys, warnings, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import encodec_pytorch_b200 as E
from encodec_pytorch_b200 import distrib

NQ, B, T, STEPS = 8, 16, 300, 4
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)


def batches():
    g = torch.Generator().manual_seed(4321)
    return [torch.randn(B, 128, T, generator=g) for _ in range(STEPS)]


def run(shard):
    torch.manual_seed(0)
    q = E.ResidualVectorQuantizer(dimension=128, n_q=NQ, bins=1024, kmeans_init=False).to(dev).train()
    codes = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad():
            for x in batches():
                lo, hi = shard(B)
                r = q(x[lo:hi].to(dev), 75, 6.0)
                codes.append(r.codes)
    return q, codes


# 1. single-rank reference on the concatenated batch (no process group yet -> world size 1 inside the helpers)
q_ref, codes_ref = run(lambda b: (0, b))
ref = {k: v.clone() for k, v in q_ref.state_dict().items()}
# 2. the sharded job
dist.init_process_group("nccl", device_id=dev)
assert distrib.world_size() == world
distrib.sync_buffers(True)        # opt-in: EMA all-reduce + k-means broadcast (default off, like the reference)
q_sh, codes_sh = run(lambda b: distrib.shard_frames(b))
lo, hi = distrib.shard_frames(B)
worst = {}
for k, v in q_sh.state_dict().items():
    d = (v - ref[k]).abs().max().item()
    scale = ref[k].abs().max().item() + 1e-30
    kind = k.split(".")[-1]
    worst[kind] = max(worst.get(kind, 0.0), d / scale)
code_mismatch = sum(int((a != b[:, lo:hi]).sum()) for a, b in zip(codes_sh, codes_ref))
t = torch.tensor([worst.get("embed", 0), worst.get("embed_avg", 0), worst.get("cluster_size", 0), float(code_mismatch)], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "steps": STEPS, "n_q": NQ, "frames_per_step": B * T,
                      "max_rel_diff_embed": t[0].item(), "max_rel_diff_embed_avg": t[1].item(),
                      "max_rel_diff_cluster_size": t[2].item(), "code_mismatches_vs_single_rank": int(t[3].item())}))
dist.destroy_process_group()
