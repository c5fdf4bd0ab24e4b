"""dev check (GPU): SLIC stages vs the oracle, bit-exact.  Run under gpurun."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
L = C.CDLL(os.path.join(os.path.dirname(__file__), '..', 'pyimsegm_b200', 'libimsegm_b200.so'))
L.isb_last_error.restype = C.c_char_p
L.isb_slic_kmeans_workspace_bytes.restype = C.c_size_t
L.isb_connectivity_workspace_bytes.restype = C.c_size_t
vp = C.c_void_p
def P(t): return vp(t.data_ptr())
def chk(rc):
    assert rc == 0, L.isb_last_error()

def run(img, sp_size, regul, name):
    H, W = img.shape[:2]
    n_seg = int(H * W / sp_size ** 2); compact = (sp_size * regul) ** 1.5
    # oracle
    t0 = time.time()
    lo, hi = img.min(), img.max()
    im = (img - lo) / float(hi - lo) if (lo != 0. or hi != 1.) else img
    im = np.ascontiguousarray(im, dtype=np.float64)
    blur = oracle.gaussian_blur(im, 1.0)
    lab_o = oracle.rgb2lab_scaled(blur, 1.0 / compact)
    km_o, cent_o = oracle.slic_kmeans(lab_o, n_seg, 10, return_centroids=True)
    seg_size = H * W / n_seg
    mn, mx = int(0.5 * seg_size), int(3 * seg_size)
    out_o = oracle.enforce_connectivity(km_o, mn, mx)
    t_or = time.time() - t0
    # gpu
    d_img = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    code = {'uint8': 0, 'uint16': 1, 'float32': 2, 'float64': 3}[str(img.dtype)]
    lab = torch.empty((3, H, W), dtype=torch.float64, device='cuda')
    mm = torch.empty(4, dtype=torch.float64, device='cuda')
    w, r = oracle.gaussian_weights(1.0)
    st = vp(torch.cuda.current_stream().cuda_stream)
    chk(L.isb_slic_prepare(P(d_img), code, H, W, 3, w.ctypes.data_as(C.POINTER(C.c_double)), r, C.c_double(1.0 / compact), 1, P(lab), P(mm), st))
    torch.cuda.synchronize()
    lab_g = lab.cpu().numpy().transpose(1, 2, 0)
    print(name, 'lab bitexact:', np.array_equal(lab_g, lab_o), 'maxdiff', np.abs(lab_g - lab_o).max())
    seeds, ty, tx = oracle.slic_seeds(H, W, n_seg)
    n = len(seeds); step = float(max(ty, tx))
    d_seeds = torch.from_numpy(seeds).cuda()
    wsb = L.isb_slic_kmeans_workspace_bytes(H, W, n, ty, tx)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    labels = torch.empty((H, W), dtype=torch.int32, device='cuda')
    cent = torch.empty((n, 5), dtype=torch.float64, device='cuda')
    # feed the ORACLE lab to isolate the kmeans stage
    lab_in = torch.from_numpy(np.ascontiguousarray(lab_o.transpose(2, 0, 1))).cuda()
    for it in (1, 2, 10):
        km_oi, cent_oi = oracle.slic_kmeans(lab_o, n_seg, it, return_centroids=True)
        chk(L.isb_slic_kmeans(P(lab_in), H, W, P(d_seeds), n, ty, tx, C.c_double(step), it, 0, P(labels), P(cent), P(ws), C.c_size_t(wsb), st))
        torch.cuda.synchronize()
        km_g = labels.cpu().numpy()
        ce = cent.cpu().numpy()
        live = ~np.isnan(cent_oi).any(1)
        print(name, 'kmeans it=%d' % it, 'labels equal:', np.array_equal(km_g, km_oi), 'ndiff', int((km_g != km_oi).sum()),
              'centroids equal:', np.array_equal(ce[live], cent_oi[live]))
    t0 = time.time()
    for _ in range(3):
        chk(L.isb_slic_kmeans(P(lab_in), H, W, P(d_seeds), n, ty, tx, C.c_double(step), 10, 0, P(labels), P(cent), P(ws), C.c_size_t(wsb), st))
    torch.cuda.synchronize()
    t_km = (time.time() - t0) / 3
    cwsb = L.isb_connectivity_workspace_bytes(H, W)
    cws = torch.empty(cwsb, dtype=torch.uint8, device='cuda')
    out = torch.empty((H, W), dtype=torch.int32, device='cuda')
    nl = torch.zeros(1, dtype=torch.int32, device='cuda')
    km_in = torch.from_numpy(km_o.astype(np.int32)).cuda()
    chk(L.isb_enforce_connectivity(P(km_in), H, W, mn, mx, P(out), P(nl), P(cws), C.c_size_t(cwsb), st))
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        chk(L.isb_enforce_connectivity(P(km_in), H, W, mn, mx, P(out), P(nl), P(cws), C.c_size_t(cwsb), st))
    torch.cuda.synchronize()
    t_cc = (time.time() - t0) / 3
    out_g = out.cpu().numpy()
    print(name, 'connectivity equal:', np.array_equal(out_g, out_o), 'ndiff', int((out_g != out_o).sum()), 'nlabels', int(nl.item()), int(out_o.max()) + 1)
    print(name, 'time oracle %.3fs | gpu kmeans %.2f ms, connectivity %.2f ms' % (t_or, t_km * 1e3, t_cc * 1e3))
    # stress: connectivity with small max_size to exercise the oversize split
    for mn2, mx2 in ((20, 60), (100, 300)):
        o2 = oracle.enforce_connectivity(km_o, mn2, mx2)
        chk(L.isb_enforce_connectivity(P(km_in), H, W, mn2, mx2, P(out), P(nl), P(cws), C.c_size_t(cwsb), st))
        torch.cuda.synchronize()
        g2 = out.cpu().numpy()
        print(name, 'connectivity(min=%d,max=%d) equal:' % (mn2, mx2), np.array_equal(g2, o2), 'ndiff', int((g2 != o2).sum()), int(nl.item()), int(o2.max()) + 1)

np.random.seed(0)
img = np.random.random((125, 150, 3)) / 2.; img[:, :75] += 0.5
run(img, 20, 0.2, 'rand125x150')
yy, xx = np.mgrid[:512, :512]
img = np.full((512, 512, 3), 0.25); img[(yy - 256) ** 2 + (xx - 256) ** 2 < 160 ** 2] = 0.75
flat = img.copy()
img = np.clip(img + np.random.normal(0, 0.05, img.shape), 0, 1)
run(img, 25, 0.2, 'disc512')
run(flat, 25, 0.2, 'flatdisc512')
run((np.clip(img, 0, 1) * 255).astype(np.uint8), 30, 0.3, 'disc512u8')
rng = np.random.RandomState(2)
H = W = 2048
pts = rng.rand(40, 2) * H; cls = rng.randint(0, 3, 40)
yy, xx = np.mgrid[:H, :W]
d = ((yy[..., None] - pts[:, 0]) ** 2 + (xx[..., None] - pts[:, 1]) ** 2).argmin(-1) if False else None
# cheaper voronoi: coarse grid upsample
g = 64; gy, gx = np.mgrid[:H // g, :W // g] * g + g / 2
lab_c = (((gy[..., None] - pts[:, 0]) ** 2 + (gx[..., None] - pts[:, 1]) ** 2).argmin(-1))
cl = np.kron(cls[lab_c], np.ones((g, g), dtype=int))
means = np.array([0.2, 0.5, 0.8])
img = means[cl][..., None] + np.array([0.0, 0.03, -0.03])
img = np.clip(img + rng.normal(0, 0.05, img.shape), 0, 1)
run(img, 29, 0.2, 'cfg2_2048')
