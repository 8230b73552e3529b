"""Per-source-line instruction counts and stall samples for one kernel.
usage: line_profile.py <src.pkl from ncu_summary.py> <section index> <libshq.so> [top N]
Maps ncu's per-SASS-instruction rows to source lines through `nvdisasm -g` of the cubin inside the .so
(instructions are matched by position inside the function)."""
import collections, os, pickle, re, subprocess, sys, tempfile
secs = pickle.load(open(sys.argv[1], 'rb'))
sec = secs[int(sys.argv[2])]
so = os.path.abspath(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith('.cubin')]
name = sec['name']
base = re.sub(r'\(.*', '', name).split('::')[-1].split('<')[0].strip()
best = None
for c in cub:
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, c)], capture_output=True, text=True).stdout
    funcs = {}; cur = None; line = None
    for l in dis.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m: cur = m.group(1); funcs[cur] = []; line = None; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(.*?);', l)
        if m and cur: funcs[cur].append((line, m.group(1)))
    for f, ins in funcs.items():
        if base in f and len(ins) == len(sec['rows']):
            # template instances: pick the one whose opcodes match
            ok = sum(1 for (a, b) in zip(ins, sec['rows']) if a[1].split()[0].lstrip('@!P0123456789 ') [:4] == b[1].split()[0].lstrip('@!P0123456789 ')[:4])
            if best is None or ok > best[0]: best = (ok, f, ins)
if best is None: sys.exit('no function with %d instructions matching %s' % (len(sec['rows']), base))
ok, f, ins = best
h = sec['hdr']; ie = h.index('Instructions Executed'); sm = h.index('# Samples'); te = h.index('Thread Instructions Executed')
agg = collections.defaultdict(lambda: [0, 0, 0])
for (ln, op), r in zip(ins, sec['rows']):
    a = agg[ln]; a[0] += int(r[ie]); a[1] += int(r[sm]); a[2] += int(r[te])
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f, 'matched opcodes %d/%d' % (ok, len(ins)), 'warp inst', tot, 'samples', ts)
srcs = {}
for ln, a in sorted(agg.items(), key=lambda x: -x[1][0] - x[1][1] * (tot / max(ts, 1)))[:top]:
    if ln is None: print('   ?', a); continue
    if ln[0] not in srcs:
        p = os.path.join(os.path.dirname(so), 'csrc', ln[0])
        srcs[ln[0]] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[ln[0]][ln[1] - 1].strip()[:90] if ln[1] - 1 < len(srcs[ln[0]]) else ''
    print("%-18s:%4d inst %5.1f%% smp %5.1f%% thr/w %4.1f | %s" % (ln[0], ln[1], 100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2] / max(a[0], 1), text))
