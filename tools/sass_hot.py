"""Hot SASS regions of one kernel from an `ncu --page source --csv` export (pickled by the caller)."""
import pickle, sys
secs = pickle.load(open(sys.argv[1], 'rb'))
s = secs[int(sys.argv[2])]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
h = s['hdr']; ie = h.index('Instructions Executed'); sm = h.index('# Samples'); te = h.index('Thread Instructions Executed')
tot = sum(int(r[ie]) for r in s['rows']); tsm = sum(int(r[sm]) for r in s['rows'])
print(s['name'][:80], 'total warp inst', tot, 'samples', tsm)
for i, r in enumerate(s['rows']):
    n = int(r[ie])
    if n > tot * thr or int(r[sm]) > tsm * thr * 2:
        print("%5d %-64s %10d %5.1f%% thr/w %4.1f smp %5.1f%%" % (i, r[1][:64].strip(), n, 100 * n / tot, int(r[te]) / max(n, 1), 100 * int(r[sm]) / tsm))
