"""Per-phase (between CTA barriers) stall-reason breakdown of the kernel last summarised by tools/ncu_summary.py.
Usage: python tools/ncu_phase_stalls.py  (reads /tmp/sass/last_src.csv)"""
import csv, collections
rows = list(csv.reader(open('/tmp/sass/last_src.csv')))
hdr = rows[1]; data = rows[2:]
iA = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
iW = hdr.index('L1 Wavefronts Shared'); iX = hdr.index('L1 Wavefronts Shared Excessive')
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
seg = 0
acc = collections.defaultdict(collections.Counter)
tot = collections.Counter(); ins = collections.Counter(); wf = collections.Counter(); wx = collections.Counter()
for r in data:
    try: e = int(r[iE]); s = int(r[iS])
    except Exception: continue
    t = r[iA].split(); op = t[1] if t[0].startswith('@') else t[0]
    tot[seg] += s; ins[seg] += e
    try: wf[seg] += int(r[iW]); wx[seg] += int(r[iX])
    except Exception: pass
    for i, name in stall_cols:
        try: acc[seg][name] += int(r[i])
        except Exception: pass
    if op.startswith('BAR') or op.startswith('SYNCS'): seg += 1
T = sum(tot.values())
for k in sorted(tot):
    if tot[k] < 0.005 * T: continue
    top = ', '.join(f'{n} {100*v/max(1,tot[k]):.0f}%' for n, v in acc[k].most_common(7))
    print(f'seg {k}: instr {ins[k]/1e6:6.2f}M samples {100*tot[k]/T:5.1f}% smem-wavefronts {wf[k]/1e6:5.2f}M (excess {wx[k]/1e6:4.2f}M) | {top}')
