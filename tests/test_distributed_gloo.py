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
