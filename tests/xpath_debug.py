"""Debug helper: tile path vs direct path on the same input (python tests/xpath_debug.py nx nsub [tile_nodes])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
from nextsim_b200 import cases, capi
import oracle_bridge as ob

def run(c, path):
    os.environ["NSX_PATH"] = path
    s = cases.make_solvers(c)[0]
    info = (C.c_int * 8)(); capi.lib().nsx_tile_info(s.h, info, 8)
    s.explicit_solve()
    out = s.download("M_VT", "M_sigma", "M_damage", "M_UM")
    s.close()
    return out, list(info)

nx, nsub = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3:
    os.environ["NSX_TILE_NODES"] = sys.argv[3]
c = cases.make_case("3km_stable", nranks=1, dyn="bbm", nx=nx, young=False)
c.params.stop_after_substeps = nsub if nsub < 120 else 0
c.params.skip_ow_smoother = 1
(a, info), (b, _) = run(c, "tiles"), run(c, "direct")
e = {k: (max(ob.rel_l2(x, y) for x, y in zip(a[k], b[k])) if k == "M_sigma" else ob.rel_l2(a[k], b[k])) for k in a}
bad = np.nonzero(np.abs(a["M_sigma"][0] - b["M_sigma"][0]) > 1e-9 * np.abs(b["M_sigma"][0]).max())[0]
print("nx", nx, "nsub", nsub, "tiles", info[:5], {k: "%.2e" % v for k, v in e.items()}, "bad elements", bad.size, bad[:8], flush=True)
