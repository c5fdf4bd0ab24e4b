"""CPU tests of the N > 1 host logic with a world-size-2 gloo group (image sharding, feature all-gather, group model)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nb_items, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pyimsegm_b200 import parallel
    from sklearn import mixture
    assert parallel.dist_info() == (rank, world)
    mine = parallel.shard_indices(nb_items)
    rng = [np.random.RandomState(100 + i) for i in range(nb_items)]
    full = [r.random_sample((5 + 3 * i, 3)) + (i % 2) for i, r in enumerate(rng)]     # ragged blocks
    got = parallel.all_gather_blocks([full[i] for i in mine], nb_items)
    assert len(got) == nb_items and all(np.array_equal(a, b) for a, b in zip(got, full))

    def compute_features(image, dict_features, sp_size, sp_regul):                      # stand-in for the GPU stage
        return None, image.reshape(-1, 3)[:: sp_size]

    def fit_model(features, nb_classes, model_type, pca_coef, use_scaler):
        return mixture.GaussianMixture(nb_classes, random_state=0).fit(features)

    images = [r.random_sample((12, 10, 3)) + (i % 2) for i, r in enumerate(rng)]
    model, fts = parallel.estim_model_classes_group_sharded(images, 2, {'color': ['mean']}, sp_size=7, compute_features=compute_features,
                                                            fit_model=fit_model)
    np.save(os.path.join(out_dir, 'means_%d.npy' % rank), model.means_)
    np.save(os.path.join(out_dir, 'nfts_%d.npy' % rank), np.array([len(f) for f in fts]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('nb_items', [5, 2, 1])
def test_sharded_group_model_world2(tmp_path, nb_items):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, nb_items, str(tmp_path)), nprocs=2, join=True)
    m0, m1 = np.load(tmp_path / 'means_0.npy'), np.load(tmp_path / 'means_1.npy')
    assert np.array_equal(m0, m1)                      # same union of features, same seed -> same model on every rank
    n0, n1 = np.load(tmp_path / 'nfts_0.npy'), np.load(tmp_path / 'nfts_1.npy')
    assert np.array_equal(n0, n1) and len(n0) == nb_items


def test_shard_indices_cover_everything_once():
    from pyimsegm_b200.parallel import shard_indices
    for n in (0, 1, 7, 64):
        for world in (1, 2, 8):
            got = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert got == list(range(n))
    assert shard_indices(64, 3, 8) == list(range(3, 64, 8))


# ---------------------------------------------------------------------------------------------------------------------
# row-band mode (pyimsegm_b200/tiled.py): the exchange protocol on CPU.  The band worker below is a numpy stand-in for the
# device kernels (same ownership rule, same halo, same 6-word exchange record), merged by the REAL communicator class over
# gloo; the result must be the oracle's whole-image k-means, bit for bit.
# ---------------------------------------------------------------------------------------------------------------------

def _numpy_band_sweeps(lab, seeds, ty, tx, step, band, halo, comm, max_iter=10):
    import torch
    H, W, _ = lab.shape
    n = len(seeds)
    cen = np.zeros((n, 5))
    cen[:, :2] = seeds
    alive = np.ones(n, dtype=bool)
    lo, hi = band.km_lo, band.km_hi
    labels = np.zeros((hi - lo, W), dtype=np.int64)
    sw = 1.0 / (step * step)
    ys = np.arange(lo, hi, dtype=np.float64)[:, None]
    xs = np.arange(W, dtype=np.float64)[None, :]
    slab = lab[lo:hi]
    for _ in range(max_iter):
        dist = np.full((hi - lo, W), np.finfo(np.float64).max)
        for k in range(n):
            if not alive[k]:
                continue
            cy, cx = cen[k, 0], cen[k, 1]
            y0, y1 = int(max(cy - 2 * ty, 0)), int(min(cy + 2 * ty + 1, H))
            x0, x1 = int(max(cx - 2 * tx, 0)), int(min(cx + 2 * tx + 1, W))
            a, b = max(y0, lo) - lo, min(y1, hi) - lo
            if a >= b:
                continue
            t_y = cy - ys[a:b]
            t_x = cx - xs[:, x0:x1]
            dc = (t_y * t_y + t_x * t_x) * sw
            d = slab[a:b, x0:x1] - cen[k, 2:]
            dcol = d[..., 0] * d[..., 0]
            dcol = dcol + d[..., 1] * d[..., 1]
            dcol = dcol + d[..., 2] * d[..., 2]
            dc = dc + dcol
            take = dist[a:b, x0:x1] > dc
            dist[a:b, x0:x1][take] = dc[take]
            labels[a:b, x0:x1][take] = k
        xchg = np.zeros((n, 6), dtype=np.int64)
        for k in range(n):
            if not alive[k] or not (band.own_lo <= int(cen[k, 0]) < band.own_hi):
                continue
            yy, xx = np.nonzero(labels == k)              # raster order
            if len(yy) == 0:
                xchg[k, 5] = 2
                continue
            assert (yy + lo).min() >= int(cen[k, 0]) - halo and (yy + lo).max() <= int(cen[k, 0]) + halo
            sums = [np.cumsum((yy + lo).astype(np.float64))[-1], np.cumsum(xx.astype(np.float64))[-1]]
            sums += [np.cumsum(slab[yy, xx, c])[-1] for c in range(3)]     # cumsum adds one at a time, in order
            xchg[k, :5] = (np.array(sums) / float(len(yy))).view(np.int64)
            xchg[k, 5] = 1
        t = torch.from_numpy(xchg)
        comm.all_reduce(t, 'sum')
        for k in range(n):
            if xchg[k, 5] == 1:
                cen[k] = xchg[k, :5].view(np.float64)
            else:
                alive[k] = False
    return labels[band.own_lo - lo:band.own_hi - lo]


def _band_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle as orc
    from conftest import synth_regions
    from pyimsegm_b200.tiled import GroupComm, plan_bands
    comm = GroupComm()
    assert (comm.rank, comm.world) == (rank, world)
    # 1) the integer sum is an exact merge of doubles, whatever their bits
    special = np.array([-0.0, np.nan, 5e-324, -np.inf, 1.0 / 3.0, -1e308])
    mine = np.where(np.arange(6) % world == rank, special.view(np.int64), 0)
    t = torch.from_numpy(mine.copy())
    comm.all_reduce(t, 'sum')
    assert np.array_equal(t.numpy(), special.view(np.int64))
    # 2) ragged band broadcast
    full = torch.zeros(7, 3, dtype=torch.int32)
    bands = plan_bands(7, world, halo=1, radius=0)
    full[bands[rank].own_lo:bands[rank].own_hi] = rank + 1
    for b in bands:
        comm.broadcast(full[b.own_lo:b.own_hi], b.index)
    assert full[:, 0].tolist() == [1, 1, 1, 1, 2, 2, 2]
    # 3) the banded sweeps against the oracle's whole-image k-means
    img = synth_regions(72, 56, seed=41, cell=16)[0]
    lab = orc.rgb2lab_scaled(orc.gaussian_blur(img, 1.0), 1.0 / 3.0)
    n_seg = 40
    want = orc.slic_kmeans(lab, n_seg)
    seeds, ty, tx = orc.slic_seeds(72, 56, n_seg)
    halo = 2 * ty + 1
    band = plan_bands(72, world, halo, 0)[rank]
    got = _numpy_band_sweeps(lab, seeds, ty, tx, float(max(1, ty, tx)), band, halo, comm)
    assert np.array_equal(got, want[band.own_lo:band.own_hi]), 'rank %d' % rank
    np.save(os.path.join(out_dir, 'band_%d.npy' % rank), got)
    dist.barrier()
    dist.destroy_process_group()


def test_row_band_exchange_world2(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_band_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / 'band_0.npy').shape[0] + np.load(tmp_path / 'band_1.npy').shape[0] == 72


def test_plan_bands():
    from pyimsegm_b200.tiled import plan_bands
    bands = plan_bands(8192, 8, halo=59, radius=4)
    assert [(b.own_lo, b.own_hi) for b in bands] == [(i * 1024, (i + 1) * 1024) for i in range(8)]
    assert (bands[0].km_lo, bands[0].raw_lo, bands[0].km_hi, bands[0].raw_hi) == (0, 0, 1083, 1087)
    assert (bands[3].km_lo, bands[3].raw_lo, bands[3].km_hi, bands[3].raw_hi) == (3013, 3009, 4155, 4159)
    assert bands[7].raw_hi == 8192
    # without a margin the uploaded rows are the raw slab; a descriptor margin (Leung-Malik: 616 rows) widens them, clipped to the image
    assert all((b.up_lo, b.up_hi) == (b.raw_lo, b.raw_hi) for b in bands)
    wide = plan_bands(8192, 8, halo=59, radius=4, margin=616)
    assert (wide[0].up_lo, wide[0].up_hi) == (0, 1024 + 616)
    assert (wide[3].up_lo, wide[3].up_hi) == (3 * 1024 - 616, 4 * 1024 + 616)
    assert (wide[7].up_lo, wide[7].up_hi) == (7 * 1024 - 616, 8192)
    assert all((a.raw_lo, a.raw_hi, a.km_lo, a.km_hi) == (b.raw_lo, b.raw_hi, b.km_lo, b.km_hi) for a, b in zip(wide, bands))
    small = plan_bands(900, 2, halo=59, radius=4, margin=616)       # the margin reaches past both ends: every band keeps the whole image
    assert [(b.up_lo, b.up_hi) for b in small] == [(0, 900), (0, 900)]
    ragged = plan_bands(10, 3, halo=1, radius=0)
    assert [(b.own_lo, b.own_hi) for b in ragged] == [(0, 4), (4, 8), (8, 10)]
    with pytest.raises(ValueError):
        plan_bands(4, 5, 1, 0)       # more bands than rows
    with pytest.raises(ValueError):
        plan_bands(7, 5, 1, 0)       # ceil(7 / 5) = 2 rows per band leaves the last band empty
