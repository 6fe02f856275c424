"""Executed warp-instructions by SASS opcode from an ncu report: python scripts/ncu_opcodes.py report.ncu-rep [warps]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; warps = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; agg = collections.Counter(); thr = collections.Counter()
for r in rows:
    if not r: continue
    if 'Source' in r and 'Instructions Executed' in r:
        hdr = r; S = hdr.index('Source'); I = hdr.index('Instructions Executed'); T = hdr.index('Thread Instructions Executed') if 'Thread Instructions Executed' in hdr else None
        continue
    if hdr and len(r) > I:
        try: n = int(r[I])
        except ValueError: continue
        toks = r[S].split()
        if not toks: continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        op = op.split('.')[0]
        agg[op] += n
        if T is not None:
            try: thr[op] += int(r[T])
            except ValueError: pass
tot = sum(agg.values())
print("total", tot, (f"per warp {tot / warps:.1f}" if warps else ""))
for op, n in agg.most_common(40):
    extra = f" per-warp {n / warps:7.1f}" if warps else ""
    lanes = f" lanes {thr[op] / n:5.1f}" if thr[op] else ""
    print(f"{op:12s} {n:12d} {100 * n / tot:5.1f}%{extra}{lanes}")
