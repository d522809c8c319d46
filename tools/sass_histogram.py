"""Blackwell opcode histogram of libnbe_b200.so, per kernel (runs without a GPU):
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.md
Counts the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): UTCHMMA (tcgen05.mma; `.2CTA` = CTA pairs),
UTMALDG (TMA tensor loads), UBLKCP (1-D bulk copies), LDTM / STTM (TMEM loads / stores), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
FFMA2 / FMUL2 / FADD2 (packed FP32), HMMA (legacy mma.sync -- expected 0)."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'brushstroke_engine_b200', 'libnbe_b200.so')
OPS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'FFMA2', 'FMUL2', 'FADD2', 'HMMA', 'FFMA', 'LDS', 'STS', 'LDG', 'STG']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r'arch = (sm_\w+)', out)))
    kern = None
    hist = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            kern = m.group(1)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and kern:
            op = m.group(1)
            base = op.split('.')[0]
            hist[kern][base] += 1
            if base == 'UTCHMMA' and '.2CTA' in op:
                hist[kern]['UTCHMMA.2CTA'] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(hist.keys()), capture_output=True, text=True).stdout.splitlines()
    print(f'# SASS opcode histogram of libnbe_b200.so (cuobjdump -sass; arch = {", ".join(arch)})\n')
    print('| kernel | instr | ' + ' | '.join(OPS) + ' |')
    print('|---|---:|' + '---:|' * len(OPS))
    tot, rest, n_rest = collections.Counter(), collections.Counter(), 0
    for (k, c), name in zip(hist.items(), demangle):
        name = re.sub(r'\(.*$', '', name).replace('void ', '')
        if sum(c.values()) == 0:
            continue
        tot.update(c)
        if not any(c.get(o, 0) for o in OPS[:11]):                   # plain SIMT kernels: one summary row
            rest.update(c)
            n_rest += 1
            continue
        print(f'| `{name}` | {sum(c.values())} | ' + ' | '.join(str(c.get(o, 0)) for o in OPS) + ' |')
    print(f'| {n_rest} other kernels (no tcgen05 / TMA / packed-FP32 opcodes) | {sum(rest.values())} | ' + ' | '.join(str(rest.get(o, 0)) for o in OPS) + ' |')
    print(f'| **all kernels** | {sum(tot.values())} | ' + ' | '.join(str(tot.get(o, 0)) for o in OPS) + ' |')


if __name__ == '__main__':
    main()
