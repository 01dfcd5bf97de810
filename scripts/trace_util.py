"""Development aid: from a DAN_B200_STACKTRACE dump, how much of the time 0 / 1 / 2 issuers are inside an op (issue start .. commit)."""
import sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/stack_trace_1.txt"
ev = sorted(((int(a, 16), int(b)) for a, b in (l.split() for l in open(path))), key=lambda x: x[1])
iv = []
for s in (0, 1):
    st = None
    for i, t in ev:
        if (i >> 28) & 1 != s: continue
        k = (i >> 24) & 15
        if k == 1: st = t
        elif k == 2 and st is not None: iv.append((st, t, s)); st = None
t0 = min(a for a, b, s in iv) + 150000; t1 = max(b for a, b, s in iv) - 50000
pts = sorted([(a, 1) for a, b, s in iv] + [(b, -1) for a, b, s in iv])
acc = {0: 0, 1: 0, 2: 0}; n = 0; last = None
for t, d in pts:
    if last is not None and t0 <= last and t <= t1: acc[n] += t - last
    n += d; last = t
tot = sum(acc.values())
print({k: round(v / tot, 3) for k, v in acc.items()}, "window cycles", tot)
# single-issuer conv duration (ops with no overlap from the other slot)
