"""
Multi-GPU plumbing: one process per GPU (``torch.distributed``, NCCL on the GPUs / gloo in CPU tests).

The hot path shards by independent images -- exactly what the reference does with its process pool
(``imsegm/utilities/experiments.py:354-410``, used by ``imsegm/pipelines.py:139-147``) -- so the data path needs NO
collective: every rank runs SLIC -> features -> GraphCut on its own images.  The only exchange is the group model
(``estim_model_classes_group``, reference pipelines.py:113-157): per-image feature blocks [N_i, D] are all-gathered and
every rank fits the same class model on the union (same data, same seed => same model; nothing to broadcast).
"""
import numpy as np


def dist_info(group=None):
    """(rank, world_size) of the default (or given) process group, (0, 1) when torch.distributed is not initialised"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_indices(nb_items, rank=None, world_size=None):
    """indices of the items this rank owns: round-robin, like a pool handing out images one at a time"""
    if rank is None or world_size is None:
        rank, world_size = dist_info()
    return list(range(rank, nb_items, world_size))


def all_gather_blocks(blocks, nb_items, group=None, device=None):
    """ all-gather per-item 2-D float64 blocks that are sharded round-robin over the ranks

    :param list(ndarray) blocks: this rank's blocks, in the order of ``shard_indices(nb_items)``
    :param int nb_items: total number of items over all ranks
    :return list(ndarray): all ``nb_items`` blocks in item order, identical on every rank
    """
    import torch
    import torch.distributed as dist
    rank, world = dist_info(group)
    if world == 1:
        return [np.asarray(b, dtype=np.float64) for b in blocks]
    backend = dist.get_backend(group)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    mine = shard_indices(nb_items, rank, world)
    assert len(mine) == len(blocks), 'rank %d holds %d blocks, expected %d' % (rank, len(blocks), len(mine))
    per_rank = (nb_items + world - 1) // world
    # 1) shapes: [per_rank, 2] (rows, cols), -1 for padding slots
    shapes = torch.full((per_rank, 2), -1, dtype=torch.int64, device=device)
    for i, b in enumerate(blocks):
        shapes[i, 0], shapes[i, 1] = int(np.shape(b)[0]), int(np.shape(b)[1])
    all_shapes = [torch.empty_like(shapes) for _ in range(world)]
    dist.all_gather(all_shapes, shapes, group=group)
    all_shapes = torch.stack(all_shapes).cpu().numpy()             # [world, per_rank, 2]
    # 2) payload: every rank sends one flat buffer padded to the largest per-rank payload
    sizes = np.where(all_shapes[..., 0] >= 0, all_shapes[..., 0] * all_shapes[..., 1], 0)
    cap = int(sizes.sum(axis=1).max())
    mine_flat = np.zeros(max(cap, 1), dtype=np.float64)
    off = 0
    for b in blocks:
        b = np.ascontiguousarray(b, dtype=np.float64).ravel()
        mine_flat[off:off + b.size] = b
        off += b.size
    flat = torch.from_numpy(mine_flat).to(device)                  # one upload of this rank's payload ...
    gathered = torch.empty((world, flat.numel()), dtype=torch.float64, device=device)
    if backend == 'nccl':
        dist.all_gather_into_tensor(gathered, flat, group=group)
    else:
        dist.all_gather(list(gathered.unbind(0)), flat, group=group)
    host = gathered.cpu().numpy()                                  # ... and one download of everybody's
    out = [None] * nb_items
    for r in range(world):
        off = 0
        for slot, item in enumerate(shard_indices(nb_items, r, world)):
            rows, cols = int(all_shapes[r, slot, 0]), int(all_shapes[r, slot, 1])
            out[item] = host[r, off:off + rows * cols].reshape(rows, cols).copy()
            off += rows * cols
    return out


def estim_model_classes_group_sharded(list_images, nb_classes, dict_features, sp_size=30, sp_regul=0.2, use_scaler=True,
                                      pca_coef=None, model_type='GMM', compute_features=None, fit_model=None, group=None):
    """ distributed ``estim_model_classes_group`` (reference pipelines.py:113-157): every rank extracts SLIC + features
    for its round-robin share of ``list_images``, the feature blocks are all-gathered, every rank fits the model on the
    union.  ``compute_features`` / ``fit_model`` default to the GPU implementations (injectable for CPU tests).

    :return tuple(model, list(ndarray)): the class model and ALL per-image feature blocks, identical on every rank
    """
    if fit_model is None:
        from .graph_cuts import estim_class_model as fit_model
    rank, world = dist_info(group)
    mine = shard_indices(len(list_images), rank, world)
    if compute_features is None:
        from .pipelines import compute_features_batch
        local = [np.nan_to_num(f) for f in compute_features_batch([list_images[i] for i in mine], dict_features, sp_size=sp_size, sp_regul=sp_regul)]
    else:
        local = []
        for i in mine:
            _, fts = compute_features(list_images[i], dict_features, sp_size=sp_size, sp_regul=sp_regul)
            local.append(np.nan_to_num(fts))
    list_features = all_gather_blocks(local, len(list_images), group=group)
    features = np.nan_to_num(np.concatenate(tuple(list_features), axis=0))
    model = fit_model(features, nb_classes, model_type, pca_coef, use_scaler)
    return model, list_features
