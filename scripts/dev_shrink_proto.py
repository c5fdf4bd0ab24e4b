"""prototype: exact node elimination before max-flow (legal pre-pushes), shrink ratio on a bench-like problem"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bench import synth_image
from sklearn import mixture, pipeline, preprocessing
img = synth_image(2, 1024, 1024)
slic, fts = oracle.compute_color2d_superpixels_features(img, ('mean',), 29, 0.2)
model = pipeline.Pipeline([('s', preprocessing.StandardScaler()), ('m', mixture.GaussianMixture(3, random_state=0, n_init=3))]).fit(fts)
proba = model.predict_proba(fts)
edges, w = oracle.edge_weights(slic, proba, 'model')
un = oracle.unary_cost(proba); pw = oracle.pairwise_cost(1.0, 3)
wi, ui, vi = oracle.integerise(w, un, pw)
N, K = ui.shape; E = len(edges)
print('N', N, 'E', E, 'w_i range', wi.min(), wi.max(), 'V', vi[0,1])
def problem(lab, alpha):
    act = lab != alpha
    U0 = np.where(act, ui[np.arange(N), alpha], 0).astype(np.int64); U1 = np.where(act, ui[np.arange(N), lab], 0).astype(np.int64)
    a, b = edges[:,0], edges[:,1]; la, lb = lab[a], lab[b]; wk = wi.astype(np.int64)
    both = act[a] & act[b]
    A = wk*vi[alpha,alpha]; B = wk*vi[alpha,lb]; C = wk*vi[la,alpha]; D = wk*vi[la,lb]
    P = np.where(both, B + C - A - D, 0)
    np.add.at(U0, a[both], A[both]); np.add.at(U1, a[both], C[both]); np.add.at(U1, b[both], (D-C)[both])
    oa = act[a] & ~act[b]; np.add.at(U0, a[oa], A[oa]); np.add.at(U1, a[oa], C[oa])
    ob = ~act[a] & act[b]; np.add.at(U0, b[ob], A[ob]); np.add.at(U1, b[ob], (wk*vi[alpha,lb])[ob])
    m = np.minimum(U0, U1)
    return act, (U1-m)*act, (U0-m)*act, P
def shrink(act, ex, tc, P):
    a, b = edges[:,0], edges[:,1]
    res_ab = P.copy(); res_ba = np.zeros_like(P)
    alive = act.copy(); ex = ex.copy(); tc = tc.copy()
    rounds = 0
    while True:
        rounds += 1
        # node pass: push to sink
        d = np.minimum(ex, tc); ex -= d; tc -= d
        ea = alive[a] & alive[b]
        out = np.zeros(N, np.int64); inn = np.zeros(N, np.int64)
        np.add.at(out, a[ea], res_ab[ea]); np.add.at(out, b[ea], res_ba[ea])
        np.add.at(inn, b[ea], res_ab[ea]); np.add.at(inn, a[ea], res_ba[ea])
        S = alive & (ex > out)
        T = alive & ~S & (tc > inn)
        if not S.any() and not T.any(): break
        # rule S: saturate out arcs
        sa = ea & S[a]; np.add.at(ex, b[sa], res_ab[sa]); 
        sb = ea & S[b]; np.add.at(ex, a[sb], res_ba[sb])
        # rule T: incoming residual -> tails' sink capacity   (tail must stay alive: skip if tail in S this round)
        ta = ea & T[b] & ~S[a]; np.add.at(tc, a[ta], res_ab[ta])
        tb = ea & T[a] & ~S[b]; np.add.at(tc, b[tb], res_ba[tb])
        alive &= ~(S | T)
    return alive, rounds
lab = np.zeros(N, np.int32)
for alpha in (1, 2):
    act, ex, tc, P = problem(lab, alpha)
    alive, rounds = shrink(act, ex, tc, P)
    ea = alive[edges[:,0]] & alive[edges[:,1]]
    print('alpha', alpha, 'active', act.sum(), '-> alive', alive.sum(), 'arcs', ea.sum(), 'rounds', rounds, 'excess nodes', (ex>0).sum(), 'sink nodes', (tc>0).sum())
    lab = oracle.alpha_expansion_int(edges, wi, ui, vi, 1) if False else lab
