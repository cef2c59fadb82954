#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: instruction share, sample
share and the top stall reasons. Usage: ncu_lines.py dump.csv [n] [sort=samples|inst]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
key = sys.argv[3] if len(sys.argv) > 3 else "samples"
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter(), 0])
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; ie = hdr.index('Instructions Executed'); ss = hdr.index('# Samples'); te = hdr.index('Thread Instructions Executed')
        st0 = hdr.index('stall_barrier'); st1 = hdr.index('stall_wait'); continue
    if hdr is None or len(r) < len(hdr): continue
    try: ln = int(r[0]); ni = int(r[ie]); s = int(r[ss]); t = int(r[te])
    except ValueError: continue
    a = agg[(cur, ln)]; a[0] += ni; a[1] += s; a[2] = r[1][:95]; a[4] += t
    for i in range(st0, st1 + 1):
        try: a[3][hdr[i].replace('stall_', '')] += int(r[i])
        except ValueError: pass
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("total warp-inst", ti, "samples", ts)
idx = 1 if key == "samples" else 0
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:n]:
    print(f"{a[1]/ts:6.3f} smp {a[0]/ti:6.3f} inst thr/inst {a[4]/max(a[0],1):5.1f} {k[0]}:{k[1]:<4d}| {a[2]} | {a[3].most_common(2)}")
