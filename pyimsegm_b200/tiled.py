"""
Row-band mode: ONE image cut into horizontal bands, one band (or a few) per GPU -- BASELINE config 5 (a single 8192 x 8192
image over 8 GPUs), SURVEY.md section 8(e) "one huge image".

What shards and what does not (reference call stack ``imsegm/pipelines.py:46-110``):

* pixel-sized work is banded: H2D of the band, min-max / blur / rgb2lab, the 10 SLIC sweeps, the colour statistics, the
  final LUT gathers and their D2H.  Per sweep the bands exchange one buffer of 6 int64 words per cluster (the centroids as
  bit patterns; summed as integers, which is an exact merge because every cluster has exactly one owner) -- the one real
  exchange step of the path, an ``all_reduce`` over NCCL.
* superpixel-sized work is replicated: every GPU gets the whole k-means label map (band broadcasts over NVLink), runs the
  connectivity pass, the adjacency extraction, the class model and the alpha-expansion on it.  These are small next to the
  pixel work and identical on every rank, so nothing has to be sent back.

The label map is bit-identical to the single-GPU path (tests/test_gpu_tiled.py): the pixel-centric assignment does not care
how the pixels are partitioned, and the raster-order sequential centroid sums are formed by the one band that owns the
cluster, over a slab that holds every member (checked on the device; an orphan pixel outside the slab makes every rank fall
back to the whole image on its own GPU).

Several bands may live on one GPU (``bands_per_rank``); the merge between them is the same integer sum done by
``isb_combine`` -- that is also how the single-GPU tests exercise every code path of the exchange.
"""
import ctypes as C
import logging

import numpy as np

from . import _lib
from .engine import FLAG_BITS, gaussian_half_kernel, get_engine, slic_seed_grid

OP_SUM_I64, OP_MAX_I64, OP_MIN_F64, OP_MAX_F64, OP_SUM_F64 = 0, 1, 2, 3, 4


class Band(object):
    """rows of one band: owned [own_lo, own_hi), k-means slab [km_lo, km_hi) = owned +- halo, raw slab = k-means slab +-
    blur radius, uploaded rows [up_lo, up_hi) = the raw slab or owned +- ``margin`` (what a descriptor with a wide footprint
    asks for), whichever reaches further (all clipped to the image)"""

    def __init__(self, index, own_lo, own_hi, H, halo, radius, margin=0):
        self.index = index
        self.own_lo, self.own_hi = own_lo, own_hi
        self.km_lo, self.km_hi = max(own_lo - halo, 0), min(own_hi + halo, H)
        self.raw_lo, self.raw_hi = max(self.km_lo - radius, 0), min(self.km_hi + radius, H)
        self.up_lo, self.up_hi = min(self.raw_lo, max(own_lo - margin, 0)), max(self.raw_hi, min(own_hi + margin, H))

    def __repr__(self):
        return 'Band(%d: own %d:%d, slab %d:%d, raw %d:%d)' % (self.index, self.own_lo, self.own_hi, self.km_lo, self.km_hi,
                                                            self.raw_lo, self.raw_hi)


def plan_bands(H, n_bands, halo, radius, margin=0):
    """equal bands of ceil(H / n_bands) rows (the last one takes what is left); every band must own at least one row"""
    rows = -(-int(H) // int(n_bands))
    bands = []
    for b in range(n_bands):
        lo, hi = b * rows, min((b + 1) * rows, H)
        if lo >= hi:
            raise ValueError('an image of %d rows cannot be cut into %d bands of %d rows' % (H, n_bands, rows))
        bands.append(Band(b, lo, hi, H, halo, radius, margin))
    return bands


class LoopbackComm(object):
    """world of one process"""
    rank, world = 0, 1

    def all_reduce(self, t, op):
        pass

    def broadcast(self, t, src):
        pass


class GroupComm(object):
    """torch.distributed process group (NCCL on the GPUs); tensors are reduced in place"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._ops = {'sum': dist.ReduceOp.SUM, 'max': dist.ReduceOp.MAX, 'min': dist.ReduceOp.MIN}

    def all_reduce(self, t, op):
        self.dist.all_reduce(t, op=self._ops[op], group=self.group)

    def broadcast(self, t, src):
        src = src if self.group is None else self.dist.get_global_rank(self.group, src)
        self.dist.broadcast(t, src=src, group=self.group)


def default_comm(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return GroupComm(group)
    return LoopbackComm()


class TiledSuperpixels(object):
    """device-resident result of :func:`slic_tiled`"""
    shape = bands = local = d_raw = d_seg = d_n_labels = nb_bound = d_feat = d_centres = d_err = None
    fell_back = False


def _combine(lib, dst_ptr, src_ptr, n, op):
    _lib.check(lib.isb_combine(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), C.c_longlong(int(n)), int(op), _lib.stream_ptr()))


def slic_tiled(image, n_segments, compactness, sigma=1.0, max_iter=10, slic_zero=False, rescale=True, comm=None,
               bands_per_rank=1, eng=None, min_size_factor=0.5, max_size_factor=3, enforce_connectivity=True, defer_check=False,
               force_whole=False, raw_margin=0):
    """ SLIC of one host image over the bands of ``comm`` (every rank passes the same image; it uploads only its rows)

    :param ndarray image: [H, W, C] host array, C in {1, 3}, dtype uint8 / uint16 / float32 / float64
    :return TiledSuperpixels: ``d_seg`` = the whole label map on this GPU (identical on every rank), ``d_raw[i]`` = the raw
        image rows ``bands[local[i]].up_lo:up_hi`` still on the device for the descriptors
    :param bool defer_check: do not synchronise to read the orphan counter ``res.d_err``; the caller reads it with its own
        results and calls again with ``force_whole=True`` when it is not zero
    :param bool force_whole: skip the banded sweeps, every rank runs them on the whole image (the fallback)
    :param int raw_margin: keep at least this many raw rows above and below the owned ones on the device (``d_raw[i]`` then holds
        the rows ``bands[local[i]].up_lo:up_hi``) -- the Leung-Malik descriptor needs its background radius + 16
    """
    eng = eng or get_engine()
    torch, lib = eng.torch, eng.lib
    comm = comm or default_comm()
    image = np.asarray(image)
    if image.ndim == 2:
        image = image[:, :, None]
    H, W, Cn = int(image.shape[0]), int(image.shape[1]), int(image.shape[2])
    code = _lib.DTYPE_CODES[str(image.dtype)]
    itemsize = image.dtype.itemsize
    st = _lib.stream_ptr()
    if sigma > 0:
        w_half, radius = gaussian_half_kernel(sigma)
    else:
        w_half, radius = np.ones(1), 0
    seeds, ty, tx = slic_seed_grid(H, W, n_segments)
    n_seeds = len(seeds)
    step = float(max(1, ty, tx))
    halo = 2 * ty + 1
    n_bands = comm.world * int(bands_per_rank)
    bands = plan_bands(H, n_bands, halo, radius, int(raw_margin))
    local = list(range(comm.rank * bands_per_rank, (comm.rank + 1) * bands_per_rank))
    owner = lambda b: b // bands_per_rank  # noqa: E731

    res = TiledSuperpixels()
    res.shape, res.bands, res.local = (H, W), bands, local
    d_seeds = eng.to_device(seeds, 'seeds')
    mm = eng.buf('tb_minmax', (4,), torch.float64)
    mm_b = eng.buf('tb_minmax_b', (4,), torch.float64)
    wsb = lib.isb_slic_kmeans_workspace_bytes(H, W, n_seeds, ty, tx)

    # 1) upload the raw rows, extrema of the owned rows
    res.d_raw = []
    for i, b in enumerate(local):
        bd = bands[b]
        raw = eng.to_device(image[bd.up_lo:bd.up_hi], 'tb%d_raw' % i)
        res.d_raw.append(raw)
        if rescale:
            own_ptr = raw.data_ptr() + (bd.own_lo - bd.up_lo) * W * Cn * itemsize
            tgt = mm if i == 0 else mm_b
            _lib.check(lib.isb_image_minmax(C.c_void_p(own_ptr), code, C.c_longlong((bd.own_hi - bd.own_lo) * W * Cn), _lib.ptr(tgt), st))
            if i > 0:
                _combine(lib, mm.data_ptr(), mm_b.data_ptr(), 1, OP_MIN_F64)
                _combine(lib, mm.data_ptr() + 8, mm_b.data_ptr() + 8, 1, OP_MAX_F64)
    if rescale:
        comm.all_reduce(mm[0:1], 'min')
        comm.all_reduce(mm[1:2], 'max')

    # 2) blur + rgb2lab of every raw slab, band descriptors
    descs, keep = [], []
    xchg = [eng.buf('tb%d_xchg' % i, (6 * n_seeds + 1,), torch.int64) for i in range(len(local))]
    mdc = [eng.buf('tb%d_maxdc' % i, (n_seeds,), torch.int64) for i in range(len(local))] if slic_zero else [None] * len(local)
    err = eng.buf('tb_err', (1,), torch.int64)
    err.zero_()
    for i, b in enumerate([] if force_whole else local):
        bd = bands[b]
        hraw = bd.raw_hi - bd.raw_lo
        lab = eng.buf('tb%d_lab' % i, (3, hraw, W), torch.float64)
        raw_ptr = C.c_void_p(res.d_raw[i].data_ptr() + (bd.raw_lo - bd.up_lo) * W * Cn * itemsize)
        _lib.check(lib.isb_slic_prepare(raw_ptr, code, hraw, W, Cn, w_half.ctypes.data_as(C.POINTER(C.c_double)), radius,
                                        C.c_double(1.0 / compactness), 2 if rescale else 0, _lib.ptr(lab), _lib.ptr(mm), st))
        slab_rows = bd.km_hi - bd.km_lo
        labels = eng.buf('tb%d_labels' % i, (slab_rows, W), torch.int32)
        ws = eng.buf('tb%d_ws' % i, (wsb,), torch.uint8)
        d = _lib.SlicBand(slab_rows=slab_rows, width=W, image_rows=H, y_off=bd.km_lo, own_lo=bd.own_lo, own_hi=bd.own_hi, halo=halo,
                          n_seeds=n_seeds, step_y=ty, step_x=tx, slic_zero=int(bool(slic_zero)), step=step,
                          lab_slab=lab.data_ptr() + (bd.km_lo - bd.raw_lo) * W * 8, plane_stride=hraw * W,
                          seeds_yx=d_seeds.data_ptr(), labels_slab=labels.data_ptr(), ws=ws.data_ptr(), ws_bytes=wsb)
        descs.append(d)
        keep.append((lab, labels, ws))
        _lib.check(lib.isb_slic_band_begin(C.byref(d), st))

    # 3) the sweeps: assign, sum the owned clusters, merge, take the merged centroids
    for _ in range(0 if force_whole else int(max_iter)):
        for i, d in enumerate(descs):
            _lib.check(lib.isb_slic_band_assign(C.byref(d), st))
            _lib.check(lib.isb_slic_band_update(C.byref(d), _lib.ptr(xchg[i]), st))
            if i > 0:
                _combine(lib, xchg[0].data_ptr(), xchg[i].data_ptr(), 6 * n_seeds + 1, OP_SUM_I64)
        comm.all_reduce(xchg[0], 'sum')
        _combine(lib, err.data_ptr(), xchg[0].data_ptr() + 8 * 6 * n_seeds, 1, OP_SUM_I64)
        for i, d in enumerate(descs):
            _lib.check(lib.isb_slic_band_import(C.byref(d), _lib.ptr(xchg[0]), _lib.ptr(mdc[i]), st))
            if slic_zero and i > 0:
                _combine(lib, mdc[0].data_ptr(), mdc[i].data_ptr(), n_seeds, OP_MAX_I64)
        if slic_zero:
            comm.all_reduce(mdc[0], 'max')
        for d in descs:
            _lib.check(lib.isb_slic_band_finalize(C.byref(d), _lib.ptr(mdc[0]), st))

    # 4) the whole k-means label map on every GPU
    full = eng.buf('tb_full', (H, W), torch.int32)
    res.d_err = err
    if not force_whole:
        for i, b in enumerate(local):
            bd = bands[b]
            full[bd.own_lo:bd.own_hi].copy_(keep[i][1][bd.own_lo - bd.km_lo:bd.own_hi - bd.km_lo])
        if comm.world > 1:
            for bd in bands:
                comm.broadcast(full[bd.own_lo:bd.own_hi], owner(bd.index))
    if force_whole or (not defer_check and int(eng.to_host(err)[0]) != 0):
        # some pixel kept the label of a cluster centred beyond the halo (no window covers it -- only degenerate inputs do
        # that): the banded sums are not trustworthy, every rank redoes the sweeps on the whole image on its own GPU
        logging.warning('slic_tiled: orphan pixels beyond the halo, redoing the sweeps on the whole image on every GPU')
        res.fell_back = True
        d_img = eng.to_device(image if Cn == 3 else image[:, :, 0], 'image')
        km, _ = eng.slic(d_img, n_segments, compactness, sigma=sigma, max_iter=max_iter, enforce_connectivity=False,
                         slic_zero=slic_zero, rescale=rescale)
        full.copy_(km)
    if not enforce_connectivity:
        res.d_seg = full
        return res
    segment_size = 1 * H * W / n_segments
    min_size, max_size = int(min_size_factor * segment_size), int(max_size_factor * segment_size)
    cwsb = lib.isb_connectivity_workspace_bytes(H, W)
    cws = eng.buf('ws_conn', (cwsb,), torch.uint8)
    out = eng.buf('labels', (H, W), torch.int32)
    n_labels = eng.buf('n_labels', (1,), torch.int32)
    _lib.check(lib.isb_enforce_connectivity(_lib.ptr(full), H, W, min_size, max_size, _lib.ptr(out), _lib.ptr(n_labels), _lib.ptr(cws),
                                            C.c_size_t(cwsb), st))
    res.d_seg, res.d_n_labels = out, n_labels
    res.nb_bound = eng.slic_label_bound(H, W, n_segments, min_size_factor)
    return res


def color_stats_tiled(res, image_dtype, channels, flags, comm=None, eng=None, feat=None, col0=0):
    """colour statistics + centroids of the banded image over ``res.d_seg``: every band accumulates its owned rows, the
    accumulators are summed over the GPUs, every GPU finishes the same [nb, 3*len(flags)] table"""
    eng = eng or get_engine()
    torch, lib = eng.torch, eng.lib
    comm = comm or default_comm()
    H, W = res.shape
    if channels != 3:
        raise ValueError('the colour statistics need a 3-channel image')
    code = _lib.DTYPE_CODES[str(np.dtype(image_dtype))]
    itemsize = np.dtype(image_dtype).itemsize
    nb = int(res.nb_bound)
    st = _lib.stream_ptr()
    bits = 0
    for f in flags:
        bits |= FLAG_BITS[f]
    acc = eng.buf('tb_acc', (nb, 6), torch.float64)
    iacc = eng.buf('tb_iacc', (nb, 3), torch.int64)
    acc.zero_()
    iacc.zero_()

    def rows(i, b):
        bd = res.bands[b]
        img_ptr = res.d_raw[i].data_ptr() + (bd.own_lo - bd.up_lo) * W * 3 * itemsize
        seg_ptr = res.d_seg.data_ptr() + bd.own_lo * W * 4
        return bd, C.c_void_p(img_ptr), C.c_void_p(seg_ptr)

    for i, b in enumerate(res.local):
        bd, img_ptr, seg_ptr = rows(i, b)
        _lib.check(lib.isb_segment_stats_accumulate(img_ptr, code, seg_ptr, bd.own_hi - bd.own_lo, W, bd.own_lo, nb, _lib.ptr(acc),
                                                    _lib.ptr(iacc), st))
    comm.all_reduce(acc, 'sum')
    comm.all_reduce(iacc, 'sum')
    var = None
    if bits & 2:
        var = eng.buf('tb_var', (nb, 3), torch.float64)
        meanf = eng.buf('tb_meanf', (nb, 3), torch.float32)
        var.zero_()
        for i, b in enumerate(res.local):
            bd, img_ptr, seg_ptr = rows(i, b)
            _lib.check(lib.isb_segment_stats_deviation(img_ptr, code, seg_ptr, bd.own_hi - bd.own_lo, W, nb, _lib.ptr(acc), _lib.ptr(iacc),
                                                       _lib.ptr(meanf), _lib.ptr(var), st))
        comm.all_reduce(var, 'sum')
    ncol = 3 * bin(bits).count('1')
    if feat is None:
        feat = eng.buf('feat', (nb, max(ncol, 1)), torch.float64)
    centres = eng.buf('centres', (nb, 2), torch.float64)
    _lib.check(lib.isb_segment_stats_finish(nb, bits, _lib.ptr(acc), _lib.ptr(var), _lib.ptr(iacc), _lib.ptr(feat), int(feat.shape[1]),
                                            int(col0), _lib.ptr(centres), None, st))
    res.d_feat, res.d_centres = feat, centres
    return feat, centres


LM_ROW_MARGIN = 616     # rows of the Leung-Malik descriptor's footprint: sigma-150 background (radius 600) + half a 33 x 33 kernel


def texture_stats_tiled(res, image_dtype, flags, bank_type='normal', comm=None, eng=None, feat=None, col0=0):
    """Leung-Malik texture statistics (reference descriptors.py:1041-1106) of the banded image over ``res.d_seg``: every band runs
    the background subtraction and the filter-bank contraction on its rows + ``LM_ROW_MARGIN`` rows of halo (``slic_tiled`` must
    have been called with ``raw_margin=LM_ROW_MARGIN``) and accumulates the sums of the rows it owns; the accumulators -- the
    per-superpixel sums and the per-battery response norms of the WHOLE image -- are summed over the GPUs, every GPU finishes
    the same [nb, n_batteries * 3 * len(flags)] block of ``feat``"""
    from .texture import _device_bank, background_kernel
    eng = eng or get_engine()
    torch, lib = eng.torch, eng.lib
    comm = comm or default_comm()
    H, W = res.shape
    code = _lib.DTYPE_CODES[str(np.dtype(image_dtype))]
    itemsize = np.dtype(image_dtype).itemsize
    nb = int(res.nb_bound)
    st = _lib.stream_ptr()
    bits = 0
    for f in flags:
        bits |= FLAG_BITS[f]
    _, d_w, NP, orient, n_batt = _device_bank(bank_type)
    w_bg, radius, mix = background_kernel()
    d_wbg = eng.const_device(w_bg, 'lm_bg_w')
    acc = eng.buf('tb_lm_acc', (int(lib.isb_lm_acc_doubles(nb, n_batt)),), torch.float64)
    counts = eng.buf('tb_lm_counts', (nb,), torch.int32)
    acc.zero_()
    counts.zero_()
    for i, b in enumerate(res.local):
        bd = res.bands[b]
        lo, hi = max(bd.own_lo - LM_ROW_MARGIN, 0), min(bd.own_hi + LM_ROW_MARGIN, H)
        if lo < bd.up_lo or hi > bd.up_hi:
            raise ValueError('the band keeps the raw rows %d:%d, the texture descriptor needs %d:%d (slic_tiled raw_margin)'
                             % (bd.up_lo, bd.up_hi, lo, hi))
        img_ptr = C.c_void_p(res.d_raw[i].data_ptr() + (lo - bd.up_lo) * W * 3 * itemsize)
        seg_ptr = C.c_void_p(res.d_seg.data_ptr() + lo * W * 4)
        wsb = lib.isb_lm_workspace_bytes(hi - lo, W, nb, n_batt)
        ws = eng.buf('ws_lm', (wsb,), torch.uint8)
        _lib.check(lib.isb_lm_texture_accumulate(img_ptr, code, seg_ptr, hi - lo, W, bd.own_lo - lo, bd.own_hi - lo, nb, _lib.ptr(d_wbg), radius,
                                                 mix.ctypes.data_as(C.POINTER(C.c_double)), _lib.ptr(d_w), NP, orient, n_batt,
                                                 _lib.ptr(acc), _lib.ptr(counts), _lib.ptr(ws), C.c_size_t(wsb), st))
    comm.all_reduce(acc, 'sum')
    comm.all_reduce(counts, 'sum')
    ncol = n_batt * 3 * bin(bits).count('1')
    if feat is None:
        feat = eng.buf('feat_lm', (nb, ncol), torch.float64)
    _lib.check(lib.isb_lm_texture_finish(nb, n_batt, bits, _lib.ptr(acc), _lib.ptr(counts), _lib.ptr(feat), int(feat.shape[1]), int(col0), st))
    return feat


def features_tiled(res, image_dtype, channels, layout, ncol, comm=None, eng=None):
    """the [nb, ncol] feature table of ``native_feature_layout`` over the banded image (``res.d_feat``) + the centroids"""
    eng = eng or get_engine()
    feat = eng.buf('feat', (int(res.nb_bound), max(ncol, 1)), eng.torch.float64)
    centres = None
    for key, flags, col0, _ in layout:
        if key == 'color':
            _, centres = color_stats_tiled(res, image_dtype, channels, flags, comm=comm, eng=eng, feat=feat, col0=col0)
        else:
            texture_stats_tiled(res, image_dtype, flags, 'short' if key.endswith('_short') else 'normal', comm=comm, eng=eng, feat=feat,
                                col0=col0)
    if centres is None:
        _, centres = color_stats_tiled(res, image_dtype, channels, (), comm=comm, eng=eng, feat=eng.buf('feat_none', (int(res.nb_bound), 1),
                                                                                                         eng.torch.float64))
    res.d_feat, res.d_centres = feat, centres
    return feat, centres


def pipe_color2d_slic_features_model_graphcut_tiled(image, nb_classes, dict_features=None, sp_size=30, sp_regul=0.2, use_scaler=True,
                                                    gc_regul=1., gc_edge_type='model', max_iter=99, comm=None, bands_per_rank=1,
                                                    want_soft=True, gather_segm=False):
    """ ``pipe_color2d_slic_features_model_graphcut`` (reference pipelines.py:46-110) for one image banded over the GPUs of
    ``comm``.  Every rank passes the same host image and gets the rows it owns:

    :return tuple: (segm [rows, W] int32, segm_soft [rows, W, K] float64 or None, (row_lo, row_hi)); with ``gather_segm``
        ``segm`` is the whole [H, W] map on every rank (``segm_soft`` stays banded: it is 8*K bytes per pixel)
    """
    from . import graph_cuts
    from .descriptors import flags_are_native, native_feature_layout
    from .graph_cuts import compute_pairwise_cost
    from .pipelines import EDGE_CAP_PER_NODE, _edge_mode
    from .superpixels import _as_rgb_like, _supported_dtype, slic_params
    if sp_regul <= 0.:
        raise ValueError('slic. regularisation must be positive')
    dict_features = {'color': ['mean']} if dict_features is None else dict_features
    layout, ncol = native_feature_layout(dict_features)
    if not layout or not flags_are_native(dict_features):
        raise NotImplementedError('the banded path computes mean / std / energy of the colours and of the Leung-Malik responses (got %r)'
                                  % dict_features)
    margin = LM_ROW_MARGIN if any(k.startswith('tLM') for k, _, _, _ in layout) else 0
    eng = get_engine()
    torch, lib = eng.torch, eng.lib
    comm = comm or default_comm()
    image = _supported_dtype(_as_rgb_like(np.asarray(image)))
    H, W = int(image.shape[0]), int(image.shape[1])
    n_seg, compact = slic_params((H, W), sp_size, sp_regul)
    if n_seg < 1:
        raise ValueError('superpixel size %r is larger than the image %r' % (sp_size, tuple(image.shape)))
    st = _lib.stream_ptr()
    K = int(nb_classes)
    n_init = max(1, int(np.sqrt(max_iter)))
    force_whole, redo_front = False, True
    while True:
        if redo_front:
            res = slic_tiled(image, n_seg, compact, sigma=1.0, comm=comm, bands_per_rank=bands_per_rank, eng=eng, defer_check=True,
                             force_whole=force_whole, raw_margin=margin)
            features_tiled(res, image.dtype, int(image.shape[2]), layout, ncol, comm=comm, eng=eng)
            nb = int(res.nb_bound)
            d_proba, _ = eng.gmm_fit_predict(res.d_feat, K, n_init, max_iter, use_scaler, graph_cuts.RANDOM_SEED, d_n=res.d_n_labels)
            redo_front = False
        lo, hi = res.bands[res.local[0]].own_lo, res.bands[res.local[-1]].own_hi
        rows = hi - lo
        seg_ptr = C.c_void_p(res.d_seg.data_ptr() + lo * W * 4)
        # segm_soft = proba[slic] of the owned rows needs only the class probabilities: its gather and its (large) download run on
        # a side stream while the main stream builds and cuts the graph
        h_soft = soft_done = None
        if want_soft:
            side = eng.side_stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                d_soft = eng.buf('segm_soft', (rows, W, K), torch.float64)
                _lib.check(lib.isb_gather(seg_ptr, C.c_longlong(rows * W), None, _lib.ptr(d_proba), K, None, _lib.ptr(d_soft),
                                          _lib.stream_ptr()))
                h_soft = eng.pinned_empty(d_soft.shape, d_soft.dtype)
                h_soft.copy_(d_soft, non_blocking=True)
                soft_done = torch.cuda.Event()
                soft_done.record(side)
        cap = max(64, EDGE_CAP_PER_NODE[0] * nb)
        pairwise = compute_pairwise_cost(gc_regul, (nb, K))
        d_edges, d_n_edges, cap = eng.adjacency(res.d_seg, nb, cap)
        _, _, unary_i, edge_wi, smooth_i = eng.gc_energies(d_proba, d_edges, cap, d_n_edges, res.d_centres, _edge_mode(gc_edge_type), 1.0,
                                                           pairwise, d_n_nodes=res.d_n_labels)
        d_labels, _, _ = eng.alpha_expansion(nb, K, cap, d_n_edges, d_edges, edge_wi, unary_i, smooth_i, -1, d_n_nodes=res.d_n_labels)
        # 5) LUT gather of the owned rows
        if gather_segm:
            d_full = eng.buf('segm', (H, W), torch.int32)
            d_segm = d_full[lo:hi]
        else:
            d_segm = eng.buf('segm', (rows, W), torch.int32)
        _lib.check(lib.isb_gather(seg_ptr, C.c_longlong(rows * W), _lib.ptr(d_labels), None, K, _lib.ptr(d_segm), None, st))
        if gather_segm and comm.world > 1:
            for r in range(comm.world):
                blo = res.bands[r * bands_per_rank].own_lo
                bhi = res.bands[(r + 1) * bands_per_rank - 1].own_hi
                comm.broadcast(d_full[blo:bhi], r)
        outs = [d_full if gather_segm else d_segm, d_n_edges, res.d_err]
        host = []
        for t in outs:
            h = eng.pinned_empty(t.shape, t.dtype)
            h.copy_(t, non_blocking=True)
            host.append(h)
        torch.cuda.current_stream().synchronize()
        if soft_done is not None:
            soft_done.synchronize()
        host.append(h_soft)
        if int(host[2][0]) != 0 and not force_whole:
            # orphan pixels beyond the halo (see slic_tiled): same answer on every rank, so every rank takes this branch
            logging.warning('banded SLIC met orphan pixels beyond the halo, redoing the sweeps on the whole image on every GPU')
            force_whole = redo_front = True
            continue
        if int(host[1][0]) <= cap:
            break
        EDGE_CAP_PER_NODE[0] *= 4
    return host[0].numpy(), (host[3].numpy() if want_soft else None), (lo, hi)
