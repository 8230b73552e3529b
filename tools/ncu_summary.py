"""Per-kernel summary (time, DRAM bytes, throughputs, top stalls) of an .ncu-rep; also pickles the source pages."""
import csv, pickle, subprocess, sys
csv.field_size_limit(10**9)
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
short = {'gpu__time_duration.sum': 'ms', 'dram__bytes_read.sum': 'rd', 'dram__bytes_write.sum': 'wr', 'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm%',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1%',
         'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2%', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%',
         'launch__registers_per_thread': 'regs', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%', 'smsp__inst_executed.sum': 'inst',
         'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'bankconf', 'launch__grid_size': 'grid', 'launch__occupancy_limit_shared_mem': 'occ_lim_smem',
         'launch__occupancy_limit_registers': 'occ_lim_reg'}
st = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
for i, d in enumerate(data):
    name = d[hdr.index('Kernel Name')][:48]
    print(i, name, ' '.join("%s=%s%s" % (v, d[hdr.index(k)][:9], units[hdr.index(k)][:2] if v in ('rd', 'wr', 'ms') else '') for k, v in short.items() if k in hdr))
    stalls = sorted(((float(d[hdr.index(s)]), s.split('stalled_')[1].split('_per')[0]) for s in st), reverse=True)[:6]
    print('    stalls:', ' '.join("%s=%.2f" % (n, v) for v, n in stalls))
def _bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
if "--traffic" in sys.argv:
    import collections, json, re
    out = sys.argv[sys.argv.index("--traffic") + 1]
    acc = collections.defaultdict(list)
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    for d in data:
        name = re.split(r'[<(]', d[hdr.index('Kernel Name')])[0].split('::')[-1].replace('void ', '').strip()
        acc[name].append(_bytes(d[ir], units[ir]) + _bytes(d[iw], units[iw]))
    tr = {k: sum(v) / len(v) for k, v in acc.items()}
    if "--workload" in sys.argv:      # the workload the capture was taken on (bench.py compares per-launch figures only on the same one)
        wname = sys.argv[sys.argv.index("--workload") + 1]
        tr["_workload"] = wname
        sys.argv = [a for a in sys.argv if a not in ("--workload", wname)]
    json.dump(tr, open(out, 'w'), indent=1, sort_keys=True)
    sys.argv = [a for a in sys.argv if a not in ("--traffic", out)]
if len(sys.argv) > 2:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    secs = []; cur = None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; secs.append(cur); continue
        if r and r[0] == 'Address': cur['hdr'] = r; continue
        if cur is not None and r: cur['rows'].append(r)
    pickle.dump(secs, open(sys.argv[2], 'wb'))
    for i, s in enumerate(secs): print(i, s['name'][:60], len(s['rows']))
