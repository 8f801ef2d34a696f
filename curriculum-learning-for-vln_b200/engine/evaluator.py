"""Navigation metrics on the World's distance tables (reference: src/engine/evaluator.py:10-146,
src/utils/dtw.py:60-82): nav / oracle error, steps, length, success and oracle rates, SPL, nDTW,
SDTW.  Off the training hot path (runs every EVAL_INTERVAL epochs); kept as host code."""
import numpy as np


class Evaluation:
    error_margin = 3.0

    def __init__(self, env):
        self.env = env
        self.gt = {it["instr_id"]: it for it in env.data}

    def _dtw(self, pred, ref):
        w = self.env.world
        m = np.full((len(pred) + 1, len(ref) + 1), np.inf)
        m[0][0] = 0
        for i in range(1, len(pred) + 1):
            for j in range(1, len(ref) + 1):
                m[i][j] = float(w.distance(pred[i - 1], ref[j - 1])) + min(m[i - 1][j], m[i][j - 1], m[i - 1][j - 1])
        dtw = m[len(pred)][len(ref)]
        ndtw = float(np.exp(-dtw / (self.error_margin * len(ref))))
        return ndtw, ndtw * (float(w.distance(pred[-1], ref[-1])) <= self.error_margin)

    def score(self, results):
        env, w = self.env, self.env.world
        s = {k: [] for k in ("nav_errors", "oracle_errors", "trajectory_steps", "trajectory_lengths",
                             "success_path_length", "ndtws", "sdtws")}
        seen = set()
        for item in results:
            iid = item["instr_id"]
            if iid in seen or iid not in self.gt:
                continue
            seen.add(iid)
            gt = self.gt[iid]
            path = [env._g_of(gt["scan"], p[0]) for p in item["trajectory"]]
            start, goal = gt["path_g"][0], gt["path_g"][-1]
            assert path[0] == start, "Result trajectories should include the start position"
            d_goal = [float(w.distance(p, goal)) for p in path]
            s["nav_errors"].append(d_goal[-1])
            s["oracle_errors"].append(min(d_goal))
            s["trajectory_steps"].append(len(path) - 1)
            length = sum(float(w.distance(a, b)) for a, b in zip(path[:-1], path[1:]))
            s["trajectory_lengths"].append(length)
            ok = d_goal[-1] < self.error_margin
            d0 = float(w.distance(start, goal))
            s["success_path_length"].append(ok * d0 / max(d0, length, 1e-9))
            nd, sd = self._dtw(path, gt["path_g"])
            s["ndtws"].append(nd), s["sdtws"].append(sd)
        assert len(seen) == len(self.gt), f"missing {len(self.gt) - len(seen)} of {len(self.gt)} instruction ids"
        n = float(len(seen))
        summary = {"nav_error": float(np.average(s["nav_errors"])), "oracle_error": float(np.average(s["oracle_errors"])),
                   "steps": float(np.average(s["trajectory_steps"])), "lengths": float(np.average(s["trajectory_lengths"])),
                   "spl": float(np.average(s["success_path_length"])), "ndtw": float(np.average(s["ndtws"])),
                   "sdtw": float(np.average(s["sdtws"])),
                   "success_rate": sum(e < self.error_margin for e in s["nav_errors"]) / n,
                   "oracle_rate": sum(e < self.error_margin for e in s["oracle_errors"]) / n}
        return summary, s


def evaluate(env, results):
    return Evaluation(env).score(results)[0]
