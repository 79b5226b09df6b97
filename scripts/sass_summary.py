"""Evidence files from a build + an ncu capture (run here, no GPU needed):
  python scripts/sass_summary.py ops  > profiles/<tag>_sass_ops.txt                    SASS opcode histogram of the shipped .so (cuobjdump)
  python scripts/sass_summary.py stalls gpurun_out/<tag>_source_sass.csv > profiles/<tag>_tc_encode_stalls_by_sass.txt
                                                                                       per-instruction warp-state samples of the capture"""
import collections, csv, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "USETMAXREG", "UTCATOMSWS", "FMNMX3", "FMNMX", "SHFL", "LDG", "LDS", "STS",
        "ATOMS", "RED", "ATOMG", "BAR", "FFMA", "STL", "LDL"]


def ops():
    so = os.path.join(ROOT, "encodec_pytorch_b200", "librvq_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    print("# SASS opcode histogram of encodec_pytorch_b200/librvq_b200.so (cuobjdump -sass, sm_100a), per kernel: the tcgen05 / TMEM / TMA evidence")
    print("# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA 1-D), UTCBAR = tcgen05.commit, SYNCS = mbarrier, "
          "USETMAXREG = setmaxnreg, RED = red.global")
    name, cnt, tot = None, collections.Counter(), 0

    def flush():
        if name:
            print(f"{name}: {tot} instructions; " + ", ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush(); name, cnt, tot = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            tot += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    cnt[k] += 1
                    break
            else:
                for k in KEYS:
                    if op.startswith(k):
                        cnt[k] += 1
                        break
    flush()


def stalls(path, min_samples=8):
    rows = list(csv.reader(open(path)))
    print(f"# per-SASS-instruction warp-state samples of {rows[0][1]} at cfg2 (ncu --set full --import-source on, one launch), instructions with >= {min_samples} samples")
    hdr = rows[1]
    ia, isrc, ins, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    st = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    print("# idx samples executed top-stall second-stall  sass")
    tot = collections.Counter()
    for n, r in enumerate(rows[2:]):
        if len(r) <= ins:
            continue
        s = int(r[ins] or 0)
        for i, nm in st:
            tot[nm] += int(r[i] or 0)
        if s < min_samples:
            continue
        top = sorted(((int(r[i] or 0), nm) for i, nm in st), reverse=True)[:2]
        print(f"{n:5d} {s:6d} {r[iex]:>9s} " + " ".join(f"{nm}:{v}" for v, nm in top) + "  " + r[isrc].strip())
    print("# totals: " + ", ".join(f"{k}={v}" for k, v in tot.most_common()))


if __name__ == "__main__":
    if sys.argv[1] == "ops":
        ops()
    else:
        stalls(sys.argv[2])
