"""Per-region instruction statistics from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: ncu_source_regions.py dump.csv [block_index] [--lines]   (regions are split at WARPSYNC markers)"""
import csv, sys
path = sys.argv[1]
block = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
lines = open(path).read().split('"Kernel Name"')
text = '"Kernel Name"' + lines[1 + block]
rows = list(csv.reader(text.splitlines()))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[idx[k]].replace(',', ''))
    except Exception: return 0.0
tot = sum(f(r, 'Instructions Executed') for r in data)
totthr = sum(f(r, 'Thread Instructions Executed') for r in data)
tots = sum(f(r, '# Samples') for r in data)
print(f"sass={len(data)} warp_inst={tot/1e9:.3f}G avg_threads={totthr/max(tot,1):.2f} samples={tots:.0f}")
regions = []; cur = dict(start=0, inst=0, thr=0, samples=0, maxexec=0)
for i, r in enumerate(data):
    if 'WARPSYNC' in r[idx['Source']]:
        cur['end'] = i; regions.append(cur); cur = dict(start=i, inst=0, thr=0, samples=0, maxexec=0)
    ie = f(r, 'Instructions Executed')
    cur['inst'] += ie; cur['thr'] += f(r, 'Thread Instructions Executed'); cur['samples'] += f(r, '# Samples')
    cur['maxexec'] = max(cur['maxexec'], ie)
cur['end'] = len(data); regions.append(cur)
for c in regions:
    if c['inst'] > 0:
        print(f"[{c['start']:4d}-{c['end']:4d}] inst={c['inst']/1e9:6.3f}G ({100*c['inst']/tot:4.1f}%) avg_thr={c['thr']/c['inst']:5.1f} "
              f"samples={100*c['samples']/max(tots,1):4.1f}% entries={c['maxexec']/1e6:.1f}M inst/entry={c['inst']/max(c['maxexec'],1):.0f}")
if '--lines' in sys.argv:
    for i, r in enumerate(data):
        ie = f(r, 'Instructions Executed')
        if ie > 0.2 * max(c['maxexec'] for c in regions):
            print(f"{i:4d} {ie/1e6:7.1f} {f(r,'Avg. Threads Executed'):5.1f} {f(r,'# Samples'):6.0f}  {r[idx['Source']][:90]}")
