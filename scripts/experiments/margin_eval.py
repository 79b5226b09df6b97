"""Evaluate the certified share under different error bounds on the fitted tables of margin_sim.py."""
import torch, math
torch.manual_seed(1)
E = torch.load("/tmp/fitted.pt")["embed"]; NQ, K, D = E.shape
NF = 6000
r = torch.randn(NF, D)
h = lambda t: t.half().float()
tot = {k: 0 for k in ["cur", "pc_exactE", "pc_worstE", "pc_sorted_worstE", "pc_sorted_exactE", "ideal_pc", "ideal_true"]}
for s in range(NQ):
    c = E[s]
    cn = (c * c).sum(1)
    nrm = cn.sqrt()
    b = -2 * c; db = (b - h(b)).norm(dim=1)            # code rounding residue
    R = r.norm(dim=1); dr = (r - h(r)).norm(dim=1)
    S = (h(r).double() @ h(b).double().t() + cn.double()[None]).float()     # approx scores (fp16 operands, exact accumulate)
    st = (r.double() @ b.double().t() + cn.double()[None]).float()
    cref = nrm.max(); dbmax = db.max()
    # current: global
    delta = (2 * 1.0625 * dbmax + 144 * 1.19e-7 * 2 * cref) * R + 2 * 1.0625 * (2 * cref + dbmax) * dr * 1.001 + 2 * (2.4e-7 * cref * cref + 6e-8)
    def share_classbatch(T, thr):
        # T [NF,K] (values compared), thr [NF]; classes = k%32, batches = k//32
        flag = T <= thr[:, None]
        Tb = T.view(NF, K // 32, 32)
        nb = (Tb.min(2).values <= thr[:, None]).sum(1)
        nc = (Tb.min(1).values <= thr[:, None]).sum(1)
        return ((nb == 1) & (nc == 1)).float().mean().item(), (flag.sum(1) == 1).float().mean().item()
    m = S.min(1).values
    tot["cur"] += share_classbatch(S, m + delta)[0]
    # per-code: delta_k = a_k R + b_k E  (half-widths)
    a_k = 1.0625 * db + 144 * 1.19e-7 * nrm * 2 * 0.5
    b_k = 1.0625 * (2 * nrm + db)
    for name, Ef in (("exactE", dr * 1.001), ("worstE", R * 2 ** -11)):
        dk = a_k[None] * R[:, None] + b_k[None] * Ef[:, None] + 1e-7
        T = S - dk
        ks = T.argmin(1)
        thr = T.min(1).values + 2 * dk.gather(1, ks[:, None])[:, 0]
        tot["pc_" + name] += share_classbatch(T, thr)[0]
        if name == "exactE":
            tot["ideal_pc"] += share_classbatch(T, thr)[1]
        # sorted by g: permute codes so batches are homogeneous, thr by winner batch's max
        g = (a_k + b_k * 2 ** -11)
        perm = g.argsort()
        Tp = T[:, perm]; dkp = dk[:, perm]
        kb = Tp.argmin(1) // 32
        dbm = dkp.view(NF, K // 32, 32).max(2).values
        thr2 = Tp.min(1).values + 2 * dbm.gather(1, kb[:, None])[:, 0]
        tot["pc_sorted_" + name] += share_classbatch(Tp, thr2)[0]
    # ideal: true gap vs nothing (share where the top-2 true gap exceeds typical actual error 4 sigma)
    a = st.argmin(1)
    r = r - c[a]
    print(s, "norms q50 %.2f q94 %.2f max %.2f | R %.2f" % (nrm.median(), nrm.kthvalue(int(K * 15 / 16)).values, cref, R.mean()),
          {k: round(v / (s + 1), 4) for k, v in tot.items()}, flush=True)
