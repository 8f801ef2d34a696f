"""engine/evaluator.py vs the REAL reference Evaluation.score (container only): random-walk trajectories over a
synthetic world scored by both; every summary entry (incl. nDTW / SDTW / CLS / SPL) must agree."""
import random
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
import clvln_b200  # noqa: E402,F401
from clvln_b200.engine.evaluator import Evaluation  # noqa: E402
from clvln_b200.environ import R2RBatch, make_items, make_world  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

w = make_world(n_scans=3, seed=4)
items = make_items(w, 60, seed=4, instr_per_path=3)
src = H.install(w, {"val": items})
import src.engine.evaluator as RE  # noqa: E402

ref_eval = RE.Evaluation(["val"], data_name="R2R")
random.seed(0)
env = R2RBatch(w, items, batch_size=8, name="val")
rng = random.Random(7)
results = []
for it in items:
    g = it["path_g"][0]
    traj = [g]
    if rng.random() < 0.5:                      # follow the ground truth for a while, then wander
        traj = list(it["path_g"][:rng.randint(1, len(it["path_g"]))])
    for _ in range(rng.randint(0, 6)):
        g = traj[-1]
        traj.append(int(w.cand_vp[g, rng.randrange(int(w.n_cand[g]))]))
    s = it["scan_idx"]
    o = int(w.scan_off[s])
    results.append({"instr_id": it["instr_id"], "trajectory": [(w.vp_names[s][g - o], 0.0, 0.0) for g in traj]})
ref_sum, _ = ref_eval.score(results)
my_sum, _ = Evaluation(env).score(results)
for k, v in ref_sum.items():
    assert k in my_sum, k
    assert np.isclose(my_sum[k], float(v), rtol=2e-6, atol=1e-6), (k, my_sum[k], v)
print("evaluation: " + ", ".join(f"{k}={my_sum[k]:.4f}" for k in ref_sum))
