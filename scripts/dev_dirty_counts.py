"""How much of a SLIC sweep is redundant on the bench image?  (CPU, oracle; dev tool, not product code.)

Runs the oracle k-means with max_iter = 1..10 on the config-2 image and reports, per sweep: pixels whose label changed,
clusters whose centroid bits changed, 32x32 tiles that see at least one moved cluster (old or new window)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import bench

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = bench.synth_image(2, side, side)
n_seg = int(side * side / bench.SP_SIZE ** 2)
compact = (bench.SP_SIZE * bench.SP_REGUL) ** 1.5
lab = oracle.rgb2lab_scaled(oracle.gaussian_blur(img, 1), 1.0 / compact)
seeds, ty, tx = oracle.slic_seeds(side, side, n_seg)
prev_l, prev_c = None, None
T = 32
nty, ntx = (side + T - 1) // T, (side + T - 1) // T
for it in range(1, 11):
    l, c = oracle.slic_kmeans(lab, n_seg, it, False, True)
    if prev_l is not None:
        changed = int((l != prev_l).sum())
        moved = np.where((c.view(np.int64) != prev_c.view(np.int64)).any(1))[0]
        act = np.zeros((nty, ntx), bool)
        for cc in (c, prev_c):
            for k in moved:
                cy, cx = cc[k, 0], cc[k, 1]
                if not np.isfinite(cy):
                    continue
                y0, y1 = int(max(cy - 2 * ty, 0)), int(min(cy + 2 * ty + 1, side))
                x0, x1 = int(max(cx - 2 * tx, 0)), int(min(cx + 2 * tx + 1, side))
                act[y0 // T:(y1 - 1) // T + 1, x0 // T:(x1 - 1) // T + 1] = True
        print('sweep %2d: changed px %8d (%.2f%%)  moved clusters %5d / %d (%.1f%%)  active tiles %.1f%%'
              % (it, changed, 100. * changed / l.size, len(moved), len(c), 100. * len(moved) / len(c), 100. * act.mean()))
    prev_l, prev_c = l, c
