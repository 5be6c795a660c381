"""CPU study: Krylov acceleration of the Peaceman-Rachford iteration on the C2 meander mask."""
import sys, os, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import cases, qpsim_b200 as Q
from pr_shift_study import build, geometric, wachspress, pr

ny = nx = int(os.environ.get("N", 256))
mask = cases.meander_mask(ny, nx, pad=8, slot=4, pitch=16, gap_len=32)
edges = Q.extract_edge_segments(mask)
bcs = cases.make_bcs(edges, "short_absorbing", Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
Gx, Gy = build(mask, bcx, bcy)
n = Gx.shape[0]; I = sp.identity(n, format="csc")
f = cases.gaussian_field(mask, cx=0.4, cy=0.5, sigma=0.05, base=1e-4, amp=2e-4)[mask]
for a in (0.3, 0.75, 1.47):
    H = 0.5 * I + a * Gx; V = 0.5 * I + a * Gy; A = (H + V).tocsc()
    b = (I - a * (Gx + Gy)) @ f
    lo, hi = 0.5, 0.5 + 4 * a
    base = pr(Gx, Gy, a, list(geometric(lo, hi, 4)), b, f)
    res = {}
    for J in (1, 2):
        shifts = [np.sqrt(lo * hi)] if J == 1 else list(geometric(lo, hi, 2))
        lus = [(spl.splu((H + r * I).tocsc()), spl.splu((V + r * I).tocsc())) for r in shifts]
        def M(rv):   # J PR steps from a zero guess for the residual equation A e = rv
            e = np.zeros(n)
            for (lh, lv), r in zip(lus, shifts):
                es = lh.solve(rv - (V - r * I) @ e)
                e = lv.solve(rv - (H - r * I) @ es)
            return e
        Mop = spl.LinearOperator((n, n), matvec=M)
        its = [0]
        def cb(x): its[0] += 1
        u0 = f.copy()
        x, info = spl.gmres(A, b, x0=u0, M=Mop, rtol=1e-14, atol=0, restart=40, maxiter=40, callback=cb, callback_type="pr_norm")
        # count iterations until the max-norm residual test of the library is met
        k = 0; xs = []
        def run(kmax):
            x, _ = spl.gmres(A, b, x0=u0, M=Mop, rtol=1e-30, atol=0, restart=kmax, maxiter=1)
            return np.max(np.abs(b - A @ x)) / np.max(np.abs(x))
        need = None
        for kmax in range(1, 25):
            if run(kmax) <= 1e-12: need = kmax; break
        its2 = [0]
        x2, info2 = spl.bicgstab(A, b, x0=u0, M=Mop, rtol=1e-14, atol=0, maxiter=40, callback=lambda xk: its2.__setitem__(0, its2[0] + 1))
        needb = None
        for kmax in range(1, 25):
            xb, _ = spl.bicgstab(A, b, x0=u0, M=Mop, rtol=1e-30, atol=0, maxiter=kmax)
            if np.max(np.abs(b - A @ xb)) / np.max(np.abs(xb)) <= 1e-12: needb = kmax; break
        res[J] = dict(gmres_iters=need, sweeps_gmres=None if need is None else need * 2 * J, bicgstab_iters=needb,
                      sweeps_bicgstab=None if needb is None else needb * 4 * J)
    print(f"alpha {a}: plain PR cyclic-4 iterations {base} ({2*base} sweeps)", res, flush=True)
