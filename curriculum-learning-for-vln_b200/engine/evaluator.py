"""Navigation metrics on the World's distance tables (reference: src/engine/evaluator.py:10-146,
src/utils/dtw.py:60-82, src/utils/cls.py:62-90): nav / oracle error, steps, length, success and oracle
rates, SPL, nDTW, SDTW, CLS.  Off the training hot path (runs every EVAL_INTERVAL epochs).  Two implementations of the
same arithmetic: host code (`score`, the reference's own structure) and one batched kernel launch over the whole
split (`score_device`, csrc/eval.cu) — the trainers use the latter on a GPU.

`dtw_scores` / `cls_score` take a distance function, so they are checked against the reference's own known-answer
doctests (dtw.py:26-34, cls.py:31-39: a 3x4 grid graph) in tests/test_host_cpu.py."""
import numpy as np


def dtw_scores(dist, pred, ref, threshold=3.0):
    """(dtw, ndtw, sdtw) of DTW.__call__ (dtw.py:60-82); `dist(a, b)` = shortest-path distance."""
    m = np.full((len(pred) + 1, len(ref) + 1), np.inf)
    m[0][0] = 0
    for i in range(1, len(pred) + 1):
        for j in range(1, len(ref) + 1):
            m[i][j] = float(dist(pred[i - 1], ref[j - 1])) + min(m[i - 1][j], m[i][j - 1], m[i - 1][j - 1])
    dtw = m[len(pred)][len(ref)]
    ndtw = float(np.exp(-dtw / (threshold * len(ref))))
    return float(dtw), ndtw, ndtw * (float(dist(pred[-1], ref[-1])) <= threshold)


def cls_score(dist, prediction, reference, threshold=3.0):
    """CLS.__call__ (cls.py:62-90) with its argument order: coverage of `reference` by `prediction`, weighted by
    the length score.  (The reference's evaluator passes (predicted_path, gt_path) into these two slots,
    evaluator.py:81-82 — i.e. it measures how well the ground truth covers the prediction; kept as is.)"""
    # (float(): the table holds fp32 values; the reference's distances are Python floats, so its arithmetic is float64)
    def length(nodes):
        return float(np.sum([float(dist(a, b)) for a, b in zip(nodes[:-1], nodes[1:])]))
    coverage = np.mean([np.exp(-np.min([float(dist(u, v)) for v in prediction]) / threshold) for u in reference])
    expected = coverage * length(reference)
    score = expected / (expected + np.abs(expected - length(prediction)))
    return float(coverage * score)


class Evaluation:
    error_margin = 3.0

    def __init__(self, env):
        self.env = env
        self.gt = {it["instr_id"]: it for it in env.data}

    def _dtw(self, pred, ref):
        _, ndtw, sdtw = dtw_scores(self.env.world.distance, pred, ref, self.error_margin)
        return ndtw, sdtw

    def score(self, results):
        env, w = self.env, self.env.world
        s = {k: [] for k in ("nav_errors", "oracle_errors", "trajectory_steps", "trajectory_lengths",
                             "success_path_length", "ndtws", "sdtws", "clss")}
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
            # evaluator.py:81-82: cls_worker(predicted_path, gt['path']) -> CLS.__call__(prediction=pred, reference=gt)
            s["clss"].append(cls_score(w.distance, path, gt["path_g"], self.error_margin))
        assert len(seen) == len(self.gt), f"missing {len(self.gt) - len(seen)} of {len(self.gt)} instruction ids"
        n = float(len(seen))
        summary = {"nav_error": float(np.average(s["nav_errors"])), "oracle_error": float(np.average(s["oracle_errors"])),
                   "steps": float(np.average(s["trajectory_steps"])), "lengths": float(np.average(s["trajectory_lengths"])),
                   "spl": float(np.average(s["success_path_length"])), "ndtw": float(np.average(s["ndtws"])),
                   "sdtw": float(np.average(s["sdtws"])), "cls": float(np.average(s["clss"])),
                   "success_rate": sum(e < self.error_margin for e in s["nav_errors"]) / n,
                   "oracle_rate": sum(e < self.error_margin for e in s["oracle_errors"]) / n}
        return summary, s


    def score_device(self, results, store):
        """The same summary through ONE kernel launch (csrc/eval.cu vln_eval_paths: a thread per trajectory, float64 on
        the HBM-resident distance table) instead of the per-trajectory Python loops above; `store` = the env's
        ops.FeatureStore.  Per-trajectory values come back as a float64 [N, 8] tensor."""
        import torch
        from .. import ops
        env = self.env
        seen, preds, refs = set(), [], []
        for item in results:
            iid = item["instr_id"]
            if iid in seen or iid not in self.gt:
                continue
            seen.add(iid)
            gt = self.gt[iid]
            path = [env._g_of(gt["scan"], p[0]) for p in item["trajectory"]]
            assert path[0] == gt["path_g"][0], "Result trajectories should include the start position"
            preds.append(path)
            refs.append(list(gt["path_g"]))
        assert len(seen) == len(self.gt), f"missing {len(self.gt) - len(seen)} of {len(self.gt)} instruction ids"
        N, P, R = len(preds), max(len(p) for p in preds), max(len(r) for r in refs)
        pa, ra = np.zeros((N, P), np.int32), np.zeros((N, R), np.int32)
        for i, (p, r) in enumerate(zip(preds, refs)):
            pa[i, :len(p)] = p
            ra[i, :len(r)] = r
        dev = store.device
        to = lambda a: torch.from_numpy(a).to(dev)
        pl, rl = to(np.array([len(p) for p in preds], np.int32)), to(np.array([len(r) for r in refs], np.int32))
        pred, ref = to(pa), to(ra)
        out = torch.empty((N, 8), dtype=torch.float64, device=dev)
        ops._call("vln_eval_paths", ops._ptr(pred), ops._ptr(pl), P, ops._ptr(ref), ops._ptr(rl), R, ops._ptr(store.dist),
                  ops._ptr(store.sq_off), ops._ptr(store.vp_local), float(self.error_margin), ops._ptr(out), N, ops._stream())
        m = out.cpu().numpy()
        n = float(N)
        summary = {"nav_error": float(np.average(m[:, 0])), "oracle_error": float(np.average(m[:, 1])),
                   "steps": float(np.average(m[:, 2])), "lengths": float(np.average(m[:, 3])),
                   "spl": float(np.average(m[:, 4])), "ndtw": float(np.average(m[:, 5])), "sdtw": float(np.average(m[:, 6])),
                   "cls": float(np.average(m[:, 7])),
                   "success_rate": float((m[:, 0] < self.error_margin).sum()) / n,
                   "oracle_rate": float((m[:, 1] < self.error_margin).sum()) / n}
        return summary, m


def evaluate(env, results, store=None):
    """Summary scores of a result list; with a FeatureStore (`agent.store_of(env)`) the batched GPU kernel computes them."""
    ev = Evaluation(env)
    return (ev.score_device(results, store) if store is not None else ev.score(results))[0]
