"""CPU study: Peaceman-Rachford iteration counts on the C2 meander mask for different shift strategies."""
import sys, os, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import cases, qpsim_b200 as Q
from scipy.special import ellipk, ellipj

def build(mask, bcx, bcy):
    ny, nx = mask.shape
    idx = -np.ones(mask.shape, dtype=np.int64); idx[mask] = np.arange(mask.sum()); n = int(mask.sum())
    def G(axis, bc):
        rows, cols, vals = [], [], []
        diag = bc[mask].copy()
        for sh in (-1, 1):
            nb = np.roll(mask, sh, axis=axis).copy()
            if axis == 0:
                if sh == 1: nb[0, :] = False
                else: nb[-1, :] = False
            else:
                if sh == 1: nb[:, 0] = False
                else: nb[:, -1] = False
            link = mask & nb
            me = idx[link]; other = np.roll(idx, sh, axis=axis)[link]
            rows.append(me); cols.append(other); vals.append(-np.ones(len(me)))
            d = np.zeros(n); np.add.at(d, me, 1.0); diag += d
        rows.append(np.arange(n)); cols.append(np.arange(n)); vals.append(diag)
        return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return G(1, bcx), G(0, bcy)

def wachspress(lo, hi, J):
    kp = lo / hi; m = 1 - kp * kp; K = ellipk(m)
    u = (2 * np.arange(J) + 1) / (2 * J) * K
    sn, cn, dn, ph = ellipj(u, m)
    return hi * dn

def geometric(lo, hi, J):
    k = np.arange(1, J + 1)
    return hi * (lo / hi) ** ((2 * k - 1) / (2 * J))

def pr(Gx, Gy, a, shifts, b, u0, tol=1e-12, maxit=60, cyclic=True):
    n = Gx.shape[0]; I = sp.identity(n, format="csc")
    H = 0.5 * I + a * Gx; V = 0.5 * I + a * Gy; A = H + V
    u = u0.copy(); lus = {}
    for k in range(maxit):
        r = b - A @ u
        if np.max(np.abs(r)) <= tol * np.max(np.abs(u)): return k
        rho = shifts[k % len(shifts)] if cyclic else shifts[min(k, len(shifts) - 1)]
        if rho not in lus: lus[rho] = (spl.splu((H + rho * I).tocsc()), spl.splu((V + rho * I).tocsc()))
        lh, lv = lus[rho]
        us = lh.solve(b - (V - rho * I) @ u)
        u = lv.solve(b - (H - rho * I) @ us)
    return maxit

if __name__ == "__main__":
    ny = nx = int(os.environ.get("N", 256))
    mask = cases.meander_mask(ny, nx, pad=8, slot=4, pitch=16, gap_len=32)
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, "short_absorbing", Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
    Gx, Gy = build(mask, bcx, bcy)
    n = Gx.shape[0]
    rng = np.random.default_rng(0)
    f = cases.gaussian_field(mask, cx=0.4, cy=0.5, sigma=0.05, base=1e-4, amp=2e-4)[mask]
    gmax = 4.0
    for a in (0.05, 0.3, 0.75, 1.47):
        I = sp.identity(n, format="csc")
        b = (I - a * (Gx + Gy)) @ f
        lo, hi = 0.5, 0.5 + a * gmax
        out = {}
        for name, sh in [("geo4", geometric(lo, hi, 4)), ("geo3", geometric(lo, hi, 3)), ("geo2", geometric(lo, hi, 2)),
                         ("geo6", geometric(lo, hi, 6)), ("geo8", geometric(lo, hi, 8)),
                         ("w4", wachspress(lo, hi, 4)), ("w5", wachspress(lo, hi, 5)), ("w6", wachspress(lo, hi, 6)), ("w8", wachspress(lo, hi, 8)),
                         ("w3", wachspress(lo, hi, 3)),
                         ("single", np.array([np.sqrt(lo * hi)]))]:
            out[name] = pr(Gx, Gy, a, list(sh), b, f)
        print(f"alpha {a}: hi/lo {hi/lo:.2f}", out, flush=True)
