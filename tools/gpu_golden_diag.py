"""Developer script: golden cases on the GPU at several tolerances (status / iterations / residuals)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cases import load_golden
from conflict_rez_b200 import solver as S
for tol in (1e-2, 1e-6, 1e-7, 1e-8):
    for name in ["single_vehicle_1", "single_vehicle_2", "single_vehicle_2_free_heading", "joint_vehicle_1_2"]:
        prob, guess, gold = load_golden(name)
        sv = S.ObcaSolver(prob, S.SolveOptions(tol=tol, constr_viol_tol=tol, max_iter=500))
        r = sv.solve(guess)
        print("%.0e %-32s %-28s it %3d  obj rel %.2e  z err %.2e  cviol %.1e du %.1e co %.1e  (oracle it %d)" % (
            tol, name, r.return_status(0), r.iters[0], abs(r.obj[0] - gold["obj"]) / abs(gold["obj"]), np.abs(r.z[0] - gold["z"]).max(),
            r.cviol[0], r.dual_inf[0], r.compl_inf[0], gold["iters"]), flush=True)
        sv.close()
