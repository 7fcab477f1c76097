"""top stall locations of one kernel from `ncu -i X.ncu-rep --page source --csv`:  python tools/ncu_top_stalls.py X.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
print(rows[0][1][:120])
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ix['# Samples']]), r))
    except (ValueError, IndexError):
        pass
tot = sum(s for s, _ in data)
agg = {n: sum(int(r[ix[n]] or 0) for _, r in data) for n in stalls}
print('total samples', tot, ' instructions', len(data))
print('by reason:', ', '.join(f'{k[6:]} {100 * v / tot:.1f}%' for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
for s, r in sorted(data, key=lambda x: -x[0])[:top]:
    why = sorted(((int(r[ix[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
    print(f'{100 * s / tot:5.1f}%  {r[ix["Source"]].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}  exec={r[ix["Instructions Executed"]]}')
