"""SASS evidence: opcode histogram per kernel of libnemoflux_gpu.so (cuobjdump -sass; no GPU needed).

    python tools/sass_histogram.py > profiles/r2_sass.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'nemoflux_b200', 'libnemoflux_gpu.so')
WATCH = ['LDG.E.NA.EFL2.256', 'LDG.E.NA.128', 'LDG.E.NA.64', 'LDG.E.NA', 'LDG.E.STRONG.GPU', 'LDG', 'STG', 'UBLKCP', 'SYNCS', 'LDS',
         'DFMA', 'DMUL', 'DADD', 'DSETP', 'FSETP', 'F2F', 'ATOMG', 'RED', 'MEMBAR', 'BAR', 'SHFL', 'CCTL', 'NANOSLEEP', 'STL', 'LDL']

txt = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r'arch = (sm_\w+)', txt)))
kernels = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
names = subprocess.run(['cu++filt'] + list(kernels), capture_output=True, text=True).stdout.splitlines()


def short(n):
    n = re.sub(r'nfx::\(anonymous namespace\)::', '', n)
    n = re.sub(r'\(int\)|\(bool\)', '', n)
    return re.sub(r'\(.*', '', n).replace('void ', '')


print('# SASS evidence for libnemoflux_gpu.so (cuobjdump -sass, built with nvcc -gencode arch=compute_100a,code=sm_100a)\n')
print(f'Architectures in the fat binary: **{", ".join(arch)}** only; {len(kernels)} kernels.  Counts are static instructions.\n')
print('What to look for: `LDG.E.NA.EFL2.256` = the 256-bit (sm_100+) `ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64` loads of the '
      'u/v stream; `LDG.E.NA.128` = the 128-bit no-allocate stream loads (float32 storage, with an evict-first `createpolicy`); '
      '`UBLKCP` + `SYNCS` = 1-D bulk copies (`cp.async.bulk`) and their mbarriers in the TMA flavour of K2; **no `DFMA` in the '
      'bit-exact per-column paths** (k2_edgeflux_ldg, k2_edgeflux_tma and the K2 half of k23_fused use separate `DMUL`/`DADD`; '
      'the `DFMA`s of k23_fused and k3_integrate are the K3 sums, whose parity bar is 1e-12 of sum |w f|; those of the K1 kernels '
      'are the Newton steps inside the correctly rounded `__ddiv_rn` and the margin pre-filter, not the predicates); '
      '`F2F` only in the rare float32 infinity fall-back; `STL`/`LDL` = register spills.\n')
cols = ['LDG.256', 'LDG.128', 'LDG other', 'STG', 'UBLKCP', 'SYNCS', 'DFMA', 'DMUL', 'DADD', 'F2F', 'ATOMG/RED', 'MEMBAR', 'BAR', 'SHFL', 'STL+LDL', 'total']
print('| kernel | ' + ' | '.join(cols) + ' |')
print('|---' * (len(cols) + 1) + '|')
for (mangled, c), name in zip(kernels.items(), names):
    def cnt(pred):
        return sum(v for k, v in c.items() if pred(k))
    l256 = cnt(lambda k: k.startswith('LDG') and '.256' in k)
    l128 = cnt(lambda k: k.startswith('LDG') and '.128' in k)
    lother = cnt(lambda k: k.startswith('LDG')) - l256 - l128
    row = [l256, l128, lother, cnt(lambda k: k.startswith('STG')), cnt(lambda k: k.startswith('UBLKCP')),
           cnt(lambda k: k.startswith('SYNCS')), cnt(lambda k: k.startswith('DFMA')), cnt(lambda k: k.startswith('DMUL')),
           cnt(lambda k: k.startswith('DADD')), cnt(lambda k: k.startswith('F2F')),
           cnt(lambda k: k.startswith('ATOMG') or k.startswith('RED') or k.startswith('ATOMS')),
           cnt(lambda k: k.startswith('MEMBAR')), cnt(lambda k: k.startswith('BAR')), cnt(lambda k: k.startswith('SHFL')),
           cnt(lambda k: k.startswith('STL') or k.startswith('LDL')), sum(c.values())]
    print(f'| `{short(name)[:70]}` | ' + ' | '.join(str(x) for x in row) + ' |')
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print('\nWhole library: ' + ', '.join(f'{k} x {v}' for k, v in sorted(tot.items(), key=lambda x: -x[1])
                                     if any(k.startswith(p) for p in ('LDG', 'UBLKCP', 'SYNCS', 'UTMA', 'HMMA', 'UTC'))) + '.')
print('\nNo tensor-core instructions (`UTC*MMA`, `HMMA`) by design: no part of the path is a dense contraction (north star).')
