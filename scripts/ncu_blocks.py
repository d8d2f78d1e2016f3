"""Group the SASS view of `ncu -i X.ncu-rep --page source --csv` into straight-line blocks with their
execution counts, active threads and stall samples.  usage: ncu_blocks.py dump.csv [kernel_index]"""
import csv, sys
text = open(sys.argv[1]).read().split('"Kernel Name"')
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(('"Kernel Name"' + text[1 + k]).splitlines()))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, key):
    try: return float(r[idx[key]].replace(',', ''))
    except Exception: return 0.0
blocks = []
for i, r in enumerate(data):
    e = f(r, 'Instructions Executed'); t = f(r, 'Avg. Threads Executed'); s = f(r, '# Samples')
    if e < 1e5: continue
    if blocks and abs(blocks[-1]['e'] - e) < 0.05 * max(e, 1) and abs(blocks[-1]['t'] - t) < 1.0 and i - blocks[-1]['end'] <= 3:
        b = blocks[-1]; b['end'] = i; b['n'] += 1; b['s'] += s; b['w'] += e
    else:
        blocks.append(dict(start=i, end=i, e=e, t=t, n=1, s=s, w=e, first=r[idx['Source']].strip()[:48]))
tot = sum(b['w'] for b in blocks); tots = sum(b['s'] for b in blocks)
print(f"warp_inst {tot/1e9:.3f}G samples {tots:.0f}")
for b in blocks:
    print(f"{b['start']:4d}-{b['end']:4d} n={b['n']:3d} exec={b['e']/1e6:6.2f}M thr={b['t']:5.1f} "
          f"winst={b['w']/1e6:6.0f}M ({100*b['w']/tot:4.1f}%) samp={100*b['s']/tots:4.1f}% {b['first']}")
