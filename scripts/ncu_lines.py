"""Aggregate an ncu report's source page by CUDA source line: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; agg = {}; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Kernel Name'): continue
    if r[0] == 'Line No': hdr = r; I = hdr.index('Instructions Executed'); S = hdr.index('# Samples'); continue
    if hdr and r[0] != '' and len(r) > max(I, S) and r[2] == '-':
        try:
            key = (cur, int(r[0])); a = agg.get(key, (0, 0, ''))
            agg[key] = (a[0] + int(r[I]), a[1] + int(r[S]), r[1].strip()[:105])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp-instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:18s}:{k[1]:4d} inst {100*v[0]/tot:5.1f}% samp {100*v[1]/max(ts,1):5.1f}%  {v[2]}")
