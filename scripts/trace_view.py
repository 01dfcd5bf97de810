"""Development aid: print the per-role event timeline dumped by DAN_B200_STACKTRACE=1 (gpurun_out/stack_trace_N.txt)."""
import sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/stack_trace_0.txt"
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 53
ev = []
for line in open(path):
    a, b = line.split()
    ev.append((int(a, 16), int(b)))
ev.sort(key=lambda x: x[1])
t0 = ev[0][1]
names = {1: "ISS_START", 2: "ISS_DONE", 3: "EPI_START", 4: "EPI_END", 5: "IO_DONE", 6: "STORE_DRAINED", 7: "READ_LANDED"}
# ISS ids are 1-based op numbers, EPI ids 0-based: align to 0-based
rows = []
for i, t in ev:
    kind = (i >> 24) & 15
    op = (i & 0xFFFF) - (1 if kind in (1, 2) else 0)
    rows.append((t - t0, (i >> 28) & 1, names[kind], op))
# per slot and op: iss_start, iss_done, epi_start, epi_end
tab = {}
for t, s, k, op in rows:
    tab.setdefault((s, op), {})[k] = t
print("slot op   iss_start  iss_dur  ->epi_gap  epi_dur  ->next_gap")
for s in (0, 1):
    for op in range(lo, hi):
        d = tab.get((s, op), {})
        nx = tab.get((s, op + 1), {})
        if "ISS_START" not in d or "EPI_END" not in d:
            continue
        print(f"{s}   {op:3d} {d['ISS_START']:10d} {d['ISS_DONE'] - d['ISS_START']:8d} {d['EPI_START'] - d['ISS_DONE']:9d} {d['EPI_END'] - d['EPI_START']:8d} "
              f"{(nx.get('ISS_START', 0) - d['EPI_END']):10d}")
    print()
