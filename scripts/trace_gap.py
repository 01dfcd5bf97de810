"""Development aid: read-boundary breakdown from a DAN_B200_STACKTRACE dump (event kinds 4 = epilogue end, 6 = store drained,
7 = next read landed, 5 = pool-add done / slot handed to the issuer, 1 = issuer starts the next read's first op)."""
import sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/stack_trace_1.txt"
ev = sorted(((int(a, 16), int(b)) for a, b in (l.split() for l in open(path))), key=lambda x: x[1])
for s in (0, 1):
    seq = [((i >> 24) & 15, t) for i, t in ev if (i >> 28) & 1 == s]
    out = []
    for n, (k, t) in enumerate(seq):
        if k == 6:
            prev4 = max((tt for kk, tt in seq[:n] if kk == 4), default=t)
            nxt = {}
            for kk, tt in seq[n + 1:n + 12]:
                nxt.setdefault(kk, tt)
            out.append((t - prev4, nxt.get(7, t) - t, nxt.get(5, t) - nxt.get(7, t), nxt.get(1, t) - nxt.get(5, t)))
    print(f"slot {s}: (bott-epi end -> store drained, -> read landed, -> pool-add done, -> issuer start)")
    for o in out[2:10]:
        print("   ", o)
