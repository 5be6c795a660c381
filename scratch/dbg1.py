import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, cases, helpers
case = cases.golden_cases()[0]
want = helpers.load_golden(case['name'])
got = helpers.run_dropin(case)
orc = helpers.run_oracle(case)
g, w = got['phonons'], want['phonons']
scale = np.max(np.abs(w), axis=-1, keepdims=True); scale = np.where(scale>0, scale, 1)
err = np.abs(g-w)/scale
print("err per time:", err.max(axis=(1,2)))
t = np.unravel_index(err.argmax(), err.shape); print("argmax", t, g[t], w[t])
print("err per omega at worst time:", np.array2string(err[t[0]].max(axis=1), precision=2))
print("oracle vs golden:", (np.abs(orc['phonons']-w)/scale).max())
print("state err:", helpers.rel_err(got['state'], want['state']))
