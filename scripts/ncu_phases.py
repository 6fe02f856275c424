"""Aggregate an ncu source page of the dense kernel by phase: python scripts/ncu_phases.py report.ncu-rep"""
import csv, subprocess, io, sys, os
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cavgym_b200", "csrc", "kernels_dense.cuh")).read().split("\n")
def find(pat):
    for i, l in enumerate(src):
        if pat in l: return i + 1
    return 10 ** 9
marks = [(find('broad_entry(double'), 'broad helpers'), (find('Box<R> dense_box'), 'road share/helpers'), (find('struct DenseEnv'), 'helpers'), (find('void dense_stage_env'), 'stage'), (find('void dense_writeback_env'), 'writeback'),
         (find('void dense_reset_env'), 'reset'), (find('bool dense_pair'), 'pair fn'), (find('void dense_transition'), 'validate'),
         (find('// ---- agents, body.step'), 'per-body step'), (find('// ---- termination cascade'), 'finish'),
         (find('// dynamic vs dynamic'), 'pair sweep'), (find('// dynamic vs static'), 'statics/offroad'),
         (find('EgoFrame<R> f;'), 'ego/zones'), (find('// ---- rewards, liveness'), 'rewards'), (find('int32_t winner = -1;'), 'tail'),
         (find('StepIO<R> io_at'), 'kernel')]
def phase(k):
    f, l = k
    if f != 'kernels_dense.cuh': return f
    name = 'top'
    for ln, nm in marks:
        if l >= ln: name = nm
    return 'dense:' + name
cur = None; agg = {}; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Kernel Name'): continue
    if r[0] == 'Line No': hdr = r; I = hdr.index('Instructions Executed'); S = hdr.index('# Samples'); continue
    if hdr and r[0] != '' and len(r) > max(I, S) and r[2] == '-':
        try:
            key = (cur, int(r[0])); a = agg.get(key, (0, 0, ''))
            agg[key] = (a[0] + int(r[I]), a[1] + int(r[S]), r[1].strip()[:100])
        except ValueError:
            pass
ts = sum(v[1] for v in agg.values()); ti = sum(v[0] for v in agg.values())
ph = {}
for k, v in agg.items():
    p = phase(k); a = ph.get(p, (0, 0)); ph[p] = (a[0] + v[0], a[1] + v[1])
print("total warp-instructions", ti, "samples", ts)
for k, v in sorted(ph.items(), key=lambda kv: -kv[1][1]): print(f"{k:40s} inst {100*v[0]/ti:5.1f}% samples {100*v[1]/ts:5.1f}%")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]: print(f"{k[0]:18s}:{k[1]:4d} inst {100*v[0]/ti:5.1f}% samp {100*v[1]/ts:5.1f}%  {v[2]}")
