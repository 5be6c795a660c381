import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, cases, helpers
import qpsim_b200 as Q
from oracle import qp_oracle as O
z = helpers.load_golden("tables_and_pixels")
for force in ("0","1"):
    os.environ["QPB_FORCE_GENERIC"]=force
    for tag in "abc":
        for rec,sc in ((True,True),(True,False),(False,True)):
            n, ph = z[f"{tag}_n_in"].copy(), z[f"{tag}_ph_in"].copy()
            dE = float(z[f"{tag}_dE"])
            args = (z[f"{tag}_Kr"], z[f"{tag}_Ks"], z[f"{tag}_rho"], z[f"{tag}_idx_diff"], z[f"{tag}_idx_sum"], z[f"{tag}_sign"])
            Q.apply_collision_step_fischer_catelani_uniform(n, ph, *args, dE, 0.3, enable_recombination=rec, enable_scattering=sc)
            wn, wp = z[f"{tag}_n_in"].copy(), z[f"{tag}_ph_in"].copy()
            O.collide(wn, wp, *args, dE, 0.3, recomb=rec, scat=sc)
            en = np.abs(n-wn)/np.abs(wn).max(axis=0,keepdims=True)
            ep = np.abs(ph-wp)/np.maximum(np.abs(wp),1e-300)
            ne = n.shape[0]
            idd = z[f"{tag}_idx_diff"]; ids = z[f"{tag}_idx_sum"]
            dset = np.unique(idd); sset = np.unique(ids)
            print(f"force={force} {tag} rec={rec} sc={sc} n_err={en.max():.2e} ph_err(elementwise): diff-bins {ep[dset].max():.2e} sum-bins {ep[sset].max():.2e} argmax n {np.unravel_index(en.argmax(), en.shape)}")
