"""Host<->device copy rates of the box (pinned memory), alone and concurrently: the bound of bench.py's end-to-end leg."""
import torch
dev = torch.device("cuda", 0)
n_in, n_out = 64 * 128 * 750 * 4, 32 * 64 * 750 * 8
hi = torch.empty(n_in, dtype=torch.uint8).pin_memory(); di = torch.empty(n_in, dtype=torch.uint8, device=dev)
ho = torch.empty(n_out, dtype=torch.uint8).pin_memory(); do = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=50):
    torch.cuda.synchronize()
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
for name, h, d in (("h2d 24.6 MB alone", True, False), ("d2h 12.3 MB alone", False, True), ("both concurrently", True, True)):
    run(h, d, 5)
    ms = run(h, d)
    gb = ((n_in if h else 0) + (n_out if d else 0)) / ms / 1e6
    print(f"{name}: {ms:.3f} ms per step, {gb:.1f} GB/s")
