"""Small cases for compute-sanitizer (memcheck / racecheck / synccheck): a few sub-cycles of the toy mesh through the C ABI.

    compute-sanitizer --tool racecheck python profiles/sanitize_case.py --path tiles --nranks 1 --substeps 6
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

ap = argparse.ArgumentParser()
ap.add_argument("--path", default="tiles")
ap.add_argument("--nranks", type=int, default=1)
ap.add_argument("--substeps", type=int, default=6)
ap.add_argument("--dyn", default="bbm")
ap.add_argument("--smoother", type=int, default=1)
a = ap.parse_args()

from nextsim_b200 import capi, cases
c = cases.make_case("toy", nranks=a.nranks, dyn=a.dyn, open_east=True)
c.params.stop_after_substeps = a.substeps
c.params.skip_ow_smoother = 0 if a.smoother else 1
solvers = cases.make_solvers(c, device=0, path=a.path, use_graph=0)
if a.nranks == 1:
    solvers[0].explicit_solve()
else:
    capi.group_explicit_solve(solvers)
for s in solvers:
    s.update()
    s.update_ice_diagnostics()
    s.check_regridding(10.0)
    got = s.download("M_VT", "M_sigma", "M_damage", "M_conc")
    chk = s.check()
    print("path", s.path, "ranks", a.nranks, "n_nan", chk.n_nan, "max_speed %.4f" % chk.max_speed)
for s in solvers:
    s.close()
