import numpy as np, sys
sys.path.insert(0,'/root/repo')
import bench, oracle
from pyimsegm_b200.superpixels import slic_params
oracle.build()
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 768
sp = 29
img = bench.synth_image(1, H, W)
n_seg, compact = slic_params((H, W), sp, 0.2)
blur = oracle.gaussian_blur(img, 1.0)
lab = oracle.rgb2lab_scaled(blur, 1.0 / compact)
seeds, ty, tx = oracle.slic_seeds(H, W, n_seg)
step = float(max(1, ty, tx)); n = len(seeds)
print('H W', H, W, 'clusters', n, 'step', step)
yy, xx = np.mgrid[:H, :W]
feat = np.concatenate([yy[..., None] / step, xx[..., None] / step, lab], axis=2).reshape(-1, 5)      # metric: D^2 = sum of squares
def cfeat(c):   # centroids (y, x, c0, c1, c2) -> 5-D metric coordinates
    f = c.copy(); f[:, 0] /= step; f[:, 1] /= step; return f
cent = [np.concatenate([seeds, np.zeros((n, 3))], axis=1)]
labels = [None]
for t in range(1, 11):
    l, c = oracle.slic_kmeans(lab, n_seg, t, False, True)
    labels.append(l.ravel()); cent.append(c)
print('centroid layout check: first cluster', cent[1][0])
T = 32
tiles_y, tiles_x = (H + T - 1) // T, (W + T - 1) // T
tile_of = ((yy // T) * tiles_x + (xx // T)).ravel()
tcy = (np.arange(tiles_y) * T + T / 2.0); tcx = (np.arange(tiles_x) * T + T / 2.0)
U = L = None
for t in range(1, 11):
    c_prev, c_new = cfeat(cent[t - 1]), cfeat(cent[t])
    lab_t = labels[t]
    if U is not None:
        # bounds carried from sweep t-1, centres moved cent[t-2] -> cent[t-1]
        delta = np.sqrt(((cfeat(cent[t - 1]) - cfeat(cent[t - 2])) ** 2).sum(1))
        # max movement over the clusters near a tile (centre within 2*step + T of the tile centre, spatially)
        cy, cx = cent[t - 1][:, 0], cent[t - 1][:, 1]
        near = (np.abs(cy[None, None, :] - tcy[:, None, None]) < 2 * step + T) & (np.abs(cx[None, None, :] - tcx[None, :, None]) < 2 * step + T)
        maxd = np.where(near, delta[None, None, :], 0).max(-1).ravel()
        a = labels[t - 1]
        Un = U + delta[a]; Ln = L - maxd[tile_of]
        stable = Un < Ln
        changed = labels[t] != labels[t - 1]
        assert not (stable & changed).any(), 'a pixel declared stable changed its label'
        # per-tile count of unstable pixels -> warps of work after compaction
        cnt = np.bincount(tile_of, weights=~stable, minlength=tiles_y * tiles_x)
        print('sweep %2d: label changes %5.2f %%  unstable %5.1f %%  warps after tile compaction %5.1f %% of full  (median delta %.4f, max %.3f)'
              % (t, 100 * changed.mean(), 100 * (~stable).mean(), 100 * np.ceil(cnt / 32).sum() / (len(tile_of) / 32), np.median(delta), delta.max()))
    else:
        stable = np.zeros(H * W, bool)
    # evaluate (exactly, all clusters within reach) the unstable pixels: new U, L; stable pixels keep updated bounds
    idx = np.nonzero(~stable)[0]
    Unew = np.empty(H * W); Lnew = np.empty(H * W)
    if U is not None:
        Unew[:] = Un; Lnew[:] = Ln
    for s0 in range(0, len(idx), 20000):
        ii = idx[s0:s0 + 20000]
        d = np.sqrt(((feat[ii, None, :] - c_prev[None, :, :]) ** 2).sum(-1))
        own = d[np.arange(len(ii)), lab_t[ii]]
        d[np.arange(len(ii)), lab_t[ii]] = np.inf
        Unew[ii] = own; Lnew[ii] = d.min(1)
    U, L = Unew, Lnew
