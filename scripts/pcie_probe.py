"""Host<->device copy rates of the box (pinned memory), alone and concurrently: the bound of bench.py's end-to-end leg.
Under torchrun every rank drives its own GPU at the same time (gloo barrier between the legs) and rank 0 prints the
per-rank and the aggregate rates: this is how the host limit of the N-GPU end-to-end leg is measured."""
import os
import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("gloo")
n_in, n_out = 64 * 128 * 750 * 4, 64 * 30000      # fp32 latents in, 10-bit packed codes out (cfg2, one step)
hi = torch.empty(n_in, dtype=torch.uint8).pin_memory(); di = torch.empty(n_in, dtype=torch.uint8, device=dev)
ho = torch.empty(n_out, dtype=torch.uint8).pin_memory(); do = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=100):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, h, d in (("h2d 24.6 MB", True, False), ("d2h 1.9 MB", False, True), ("both concurrently", True, True)):
    run(h, d, 5)
    ms = run(h, d)
    gb = ((n_in if h else 0) + (n_out if d else 0)) / ms / 1e6
    t = torch.tensor([ms, gb], dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        mss = [float(v[0]) for v in allv]
        print(f"{name}, {world} rank(s) at once: per rank {min(mss):.3f}..{max(mss):.3f} ms per step, "
              f"aggregate {sum(float(v[1]) for v in allv):.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
