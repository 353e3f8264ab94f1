"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked markdown summary.
usage: python profiles/summarize.py <launch-list.csv> <title> [name.ncu-rep ...]"""
import collections
import csv
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
           'launch__block_size', 'lts__t_sector_hit_rate.pct',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0.0, 0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6}[row['Metric Unit']]
        k = row['Kernel Name'].split('(')[0]
        agg[k][0] += v
        agg[k][1] += 1
        tot += v
    print('%d launches, %.2f ms summed device time (cold-cache, serialised by ncu: compare shares).\n' % (
        sum(v[1] for v in agg.values()), tot / 1e6))
    print('| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print('| `%s` | %d | %.1f | %.2f | %.3f |' % (k[:90], v[1], v[0] / 1e3, v[0] / v[1] / 1e3, v[0] / tot))


def report(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, universal_newlines=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr = r[0]
    print('\n### `%s`\n' % r[2][hdr.index('Kernel Name')].split('(')[0])
    print('| metric | unit | ' + ' | '.join('launch %d' % (i + 1) for i in range(len(r) - 2)) + ' |')
    print('|---|---|' + '---|' * (len(r) - 2))
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print('| %s | %s | %s |' % (m, r[1][i], ' | '.join(row[i] for row in r[2:])))


if __name__ == '__main__':
    print('# %s\n' % sys.argv[2])
    launches(sys.argv[1])
    for p in sys.argv[3:]:
        report(p)
