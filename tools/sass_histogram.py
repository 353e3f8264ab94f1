"""SASS op histogram of the built library (cuobjdump -sass), written as markdown: which kernels carry the Blackwell
instructions (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG, cp.async = LDGSTS, mbarrier = SYNCS).
usage: python tools/sass_histogram.py > profiles/r2_sass_histogram.md   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(HERE, 'bilevel-graph-neural-network_b200', '_C', 'libbignn_b200.so')

WHAT = [('UTCHMMA', 'tcgen05.mma (kind::tf32)'), ('UTCBAR', 'tcgen05.commit -> mbarrier'),
        ('LDTM', 'tcgen05.ld (TMEM -> registers)'), ('STTM', 'tcgen05.st (registers -> TMEM)'),
        ('UTCATOMSWS', 'tcgen05.alloc/dealloc'), ('UTMALDG', 'TMA tensor load (cp.async.bulk.tensor)'),
        ('UTMASTG', 'TMA tensor store'), ('UBLKCP', 'TMA bulk copy'), ('LDGSTS', 'cp.async (LDGSTS)'),
        ('SYNCS', 'mbarrier arrive / try_wait / test_wait'), ('NANOSLEEP', 'nanosleep back-off'),
        ('DFMA', 'fp64 FMA (BatchNorm statistics)'), ('DADD', 'fp64 add'), ('HMMA', 'legacy mma.sync (none expected)')]
COLS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'LDGSTS', 'SYNCS', 'UTCBAR']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, universal_newlines=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(kernels), stdout=subprocess.PIPE, universal_newlines=True).stdout.splitlines()
    total = collections.Counter()
    for c in kernels.values():
        total.update(c)
    print('# SASS op histogram of %s (cuobjdump -sass, sm_100a)\n' % os.path.relpath(lib, HERE))
    print('Made by `python tools/sass_histogram.py`.  Whole library: %d kernels, %d instructions.  Blackwell-specific mnemonics:\n'
          % (len(kernels), sum(total.values())))
    print('| mnemonic | meaning | count |\n|---|---|---|')
    for k, what in WHAT:
        print('| `%s` | %s | %d |' % (k, what, total[k]))
    print('\nPer kernel (kernels that use the tensor cores, TMEM, TMA or cp.async):\n')
    print('| kernel | instructions | ' + ' | '.join(COLS) + ' |\n|---|---|' + '---|' * len(COLS))
    for (k, c), name in zip(kernels.items(), names):
        if any(c[x] for x in ('UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'LDGSTS')):
            print('| `%s` | %d | %s |' % (name.split('(')[0], sum(c.values()), ' | '.join(str(c[x]) for x in COLS)))
    print('\nTop 12 mnemonics of the fused layer kernel variants:\n')
    for (k, c), name in zip(kernels.items(), names):
        if 'k_gin_layer_fwd' in name:
            print('* `%s`: %s' % (name.split('(')[0], ', '.join('%s %d' % kv for kv in c.most_common(12))))


if __name__ == '__main__':
    main()
