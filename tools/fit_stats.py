import sys, warnings, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import load_golden
import glomargridding_b200.ellipse as E
import time, torch
g = load_golden("ellipse_fit.npz")
for c in range(int(g["n_cases"])):
    aniso, rot, pd, unit = (bool(x) for x in g[f"c{c}_meta"])
    m = E.EllipseModel(aniso, rot, pd, float(g[f"c{c}_v"]), unit_sigma=unit)
    res, se, b = m.fit(g[f"c{c}_X"], g[f"c{c}_y"].copy())
    w = g[f"c{c}_fit"]
    print(c, m.fform, 'nit', res.nit, int(w[1]), 'nfev', res.nfev, int(w[2]), 'maxrel', float(np.max(np.abs(res.x-g[f"c{c}_fit_x"])/np.abs(g[f"c{c}_fit_x"]))), 'fun rel', abs(res.fun-w[0])/abs(w[0]))
class _Coord:
    def __init__(self, values): self.values = np.asarray(values)
g = load_golden("ellipse_builder.npz")
data = np.ma.masked_array(g["data"], g["mask"])
coords = {"time": _Coord(np.arange(data.shape[0])), "latitude": _Coord(g["lats"]), "longitude": _Coord(g["lons"])}
b = E.EllipseBuilder(data, coords)
for tag in ("ard_pd","iso_deg"):
    if tag == "ard_pd":
        m = E.EllipseModel(anisotropic=True, rotated=True, physical_distance=True, v=1.5, unit_sigma=False)
        kw = dict(max_distance=3000.0, min_distance=0.3, delta_x_method="Modified_Met_Office", tol=1e-4)
    else:
        m = E.EllipseModel(anisotropic=False, rotated=False, physical_distance=False, v=0.5, unit_sigma=True)
        kw = dict(max_distance=35.0, min_distance=0.3, delta_x_method=None, tol=1e-4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.cuda.synchronize(); t0=time.perf_counter()
        params = b.compute_params(default_value=-999.0, matern_ellipse=m, **kw)
        torch.cuda.synchronize(); t1=time.perf_counter()
        params = b.compute_params(default_value=-999.0, matern_ellipse=m, **kw)
        torch.cuda.synchronize(); t2=time.perf_counter()
    want, aux = g[f"fit_{tag}"], g[f"fit_{tag}_aux"]
    names = list(m.supercategory_params.keys())
    got = np.column_stack([np.asarray(params[n])[b.yi_masked, b.xi_masked] for n in names])
    nfev = np.array([r.nfev for r in b.fit_results]); fun=np.array([r.fun for r in b.fit_results])
    same = (got[:,-1]==want[:,-1]) & (nfev==aux[:,0])
    print(tag, 'time', t1-t0, t2-t1, 'cells', len(want), 'identical path', same.sum(), 'nit same', (got[:,-1]==want[:,-1]).sum(), 'qc same', (got[:,-2]==want[:,-2]).sum())
    rel = np.abs(got[:,:-2]-want[:,:-2])/np.maximum(np.abs(want[:,:-2]),1e-12)
    print('  max rel per param', rel.max(0), 'p99', np.percentile(rel,99,axis=0), 'n>1e-6', (rel>1e-6).sum(0), 'fun rel max', np.max(np.abs(fun-aux[:,2])/np.abs(aux[:,2])))
    bad = np.flatnonzero((rel>1e-6).any(1))[:8]
    for k in bad: print('   cell',k,'got',got[k],'want',want[k], 'nfev', nfev[k], aux[k,0])
