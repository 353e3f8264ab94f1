"""Reads the pipeline time stamps written by bignn_gin_layer_fwd under BIGNN_GL_TRACE=<file> (CTA 0, first 64 tiles)
and prints, per tile, when each role reached each event (microseconds from the first event; SM clock 1.965 GHz)."""
import sys
import numpy as np
t = np.loadtxt(sys.argv[1])
mhz = float(sys.argv[2]) if len(sys.argv) > 2 else 1965.0
names = ['x_full seen', 'agg done', 'z_full', 'm1 seen', 't_full', 'm2 seen', 'acc2 read', 'z_empty', 'TMA issue', 'MMA1 issue', 'MMA2 issue', 'row it 1', 'row it 2', 'row it 3', 'prefetched']
t0 = t[t > 0].min()
rows = []
for i in range(8, 24):
    rows.append([(t[i, e] - t0) / mhz if t[i, e] > 0 else float('nan') for e in range(15)])
rows = np.asarray(rows)
print('tile ' + ' '.join('%11s' % n for n in names))
for i, r in enumerate(rows):
    print('%4d ' % (i + 8) + ' '.join('%11.2f' % v for v in r))
r = rows
d = np.diff(rows, axis=0)
print('period per tile (us): ' + ' '.join('%s %.2f' % (n, v) for n, v in zip(names, np.nanmean(d, axis=0))))
r = rows
print('producer warp detail (us): idx ready+prefetch -> x_full wait %.2f | x_full -> row iteration 1 done %.2f | it 2 %.2f | it 3 %.2f' % (np.nanmean(r[:, 0] - r[:, 14]), np.nanmean(r[:, 11] - r[:, 0]), np.nanmean(r[:, 12] - r[:, 11]), np.nanmean(r[:, 13] - r[:, 12])))
print('E2 detail (us): acc2 read -> staged (barrier) %.2f | copy-out loop %.2f | statistics + barriers %.2f' % (np.nanmean(r[:, 8] - r[:, 6]), np.nanmean(r[:, 14] - r[:, 8]), np.nanmean(r[:, 7] - r[:, 14])))
print('mean stage times (us): TMA->x_full %.2f | aggregate %.2f | lo+fence %.2f | z_full->MMA1 issue %.2f | MMA1 issue->m1 seen %.2f | E1 %.2f | t_full->MMA2 issue %.2f | MMA2 issue->m2 seen %.2f | E2 tmem+stage %.2f | E2 copy-out+stats %.2f | z_empty->next TMA %.2f' % (
    np.nanmean(r[:, 0] - r[:, 8]), np.nanmean(r[:, 1] - r[:, 0]), np.nanmean(r[:, 2] - r[:, 1]), np.nanmean(r[:, 9] - r[:, 2]),
    np.nanmean(r[:, 3] - r[:, 9]), np.nanmean(r[:, 4] - r[:, 3]), np.nanmean(r[:, 10] - r[:, 4]), np.nanmean(r[:, 5] - r[:, 10]),
    np.nanmean(r[:, 6] - r[:, 5]), np.nanmean(r[:, 7] - r[:, 6]), np.nanmean(r[2:, 8] - r[:-2, 7])))
