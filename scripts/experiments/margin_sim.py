"""CPU study: share of frame-stages the certified argmin would settle under different score-error bounds, on codebooks
fitted like bench.py's `trained_like` block (k-means init + EMA steps on N(0,1) latents).  Not part of the product."""
import torch, math, sys
torch.manual_seed(0)
D, K = 128, 1024
NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 26
sets = [torch.randn(NF, D) for _ in range(4)]

def assign(r, e):
    d = (r * r).sum(1, keepdim=True) - 2 * r @ e.t() + (e * e).sum(1)[None]
    return d.argmin(1)

def kmeans(x, iters=10):
    e = x[torch.randperm(len(x))[:K]].clone()
    for _ in range(iters):
        a = assign(x, e)
        cnt = torch.bincount(a, minlength=K).float()
        s = torch.zeros(K, D).index_add_(0, a, x)
        new = s / cnt.clamp(min=1)[:, None]
        e = torch.where(cnt[:, None] > 0, new, e)
    return e, cnt

embed = [None] * NQ; cs = [None] * NQ; avg = [None] * NQ
for step in range(STEPS):
    r = sets[step % 4].clone()
    for s in range(NQ):
        if embed[s] is None:
            embed[s], cs[s] = kmeans(r); avg[s] = embed[s] * cs[s][:, None]   # reference: embed_avg = embed clone; cluster_size = bins
            avg[s] = embed[s].clone()
        a = assign(r, embed[s])
        q = embed[s][a]
        cnt = torch.bincount(a, minlength=K).float()
        su = torch.zeros(K, D).index_add_(0, a, r)
        cs[s] = cs[s] * .99 + cnt * .01
        avg[s] = avg[s] * .99 + su * .01
        n = cs[s].sum()
        sm = (cs[s] + 1e-5) / (n + K * 1e-5) * n
        embed[s] = avg[s] / sm[:, None]
        r = r - q
    print("step", step, "resid", float(r.norm(dim=1).mean()), flush=True)
torch.save({"embed": torch.stack(embed)}, "/tmp/fitted.pt")
