import sys, random, numpy as np, torch
sys.path.insert(0, "/root/repo")
import clvln_b200
from clvln_b200.environ import make_world, make_items, world as W
from oracle import ref_harness as H, ref_loader
w = make_world(n_scans=3, seed=1)
items = make_items(w, 40, seed=1)
src = H.install(w, {"train": items})
import src.environ as environ
random.seed(2020)
tok = H.StubTokenizer(items)
fs = H.feature_store(w)
env = environ.R2RBatch(fs, batch_size=8, splits=["train"], tokenizer=tok)
H.warm_candidate_buffer(env)
obs = env.reset()
print(len(obs), obs[0]['viewIndex'], obs[0]['heading'], len(obs[0]['candidates']), obs[0]['teacher'], obs[0]['distance'], [o['instr_length'] for o in obs])
# verify against world tables
name2g = {w.long_id(g): g for g in range(w.n_vp)}
loc4 = W.static_loc4()
for ob in obs:
    g = name2g[ob['scan']+"_"+ob['viewpointId']]
    assert len(ob['candidates']) == w.n_cand[g]
    v = ob['viewIndex']
    exp = np.concatenate([w.table[g].float().numpy(), np.repeat(loc4[v], 32, axis=1)], 1)
    assert np.array_equal(ob['feature'], exp)
    for j, c in enumerate(ob['candidates']):
        assert name2g[ob['scan']+"_"+c['nextViewpointId']] == w.cand_vp[g, j]
        assert c['absViewIndex'] == w.cand_view[g, j]
        e = np.concatenate([w.table[g, c['absViewIndex']].float().numpy(), np.repeat(w.cand_ang4[g, j, v % 12], 32)])
        assert np.array_equal(c['feature'], e), (j,)
    item = [it for it in items if it['instr_id'] == ob['instr_id']][0]
    goal = item['path_g'][-1]
    assert np.float32(ob['distance']) == w.distance(g, goal), (ob['distance'], w.distance(g, goal))
    t = w.teacher_action(g, goal)
    if t < len(ob['candidates']): assert ob['candidates'][t]['nextViewpointId'] == ob['teacher']
    else: assert ob['teacher'] == ob['viewpointId']
# step with teacher actions a few times
traj = [{'instr_id': ob['instr_id'], 'path': [(ob['viewpointId'], ob['heading'], ob['elevation'])]} for ob in obs]
for step in range(8):
    acts = []
    for ob in obs:
        a = -1
        for k, c in enumerate(ob['candidates']):
            if c['nextViewpointId'] == ob['teacher']: a = k
        acts.append(a)
    obs = env.step(np.array(acts), obs, traj)
print([ob['distance'] for ob in obs])
# ---- first-visit observations.  make_candidate's first visit to a viewpoint computes a candidate's
# relative heading as (state.heading - base) + rel_heading (common_env.py:249-256); every later visit reads the buffer and
# computes (state.heading + rel_heading) - base (:283-285).  The two associations differ by at most one fp64 rounding of the
# heading, i.e. after float32(sin / cos) by at most one unit in the last place of the 128 angle features; the image half,
# the candidate order and the view indices are identical.  The product tables (and the checks above) hold the BUFFERED
# arithmetic, the one every observation but a viewpoint's very first uses.
env2 = environ.R2RBatch(fs, batch_size=8, splits=["train"], tokenizer=tok)          # cold buffer
worst, n_diff, n_tot = 0.0, 0, 0
for long_id, f in list(fs.items())[:200]:
    scan, vp = long_id.split("_", 1)
    if scan not in env2.scans:
        continue
    for view in (12, 17, 23):
        env2.buffered_state_dict.pop(long_id, None)
        first = env2.make_candidate(f, scan, vp, view)
        again = env2.make_candidate(f, scan, vp, view)
        assert [c["nextViewpointId"] for c in first] == [c["nextViewpointId"] for c in again]
        assert [c["absViewIndex"] for c in first] == [c["absViewIndex"] for c in again]
        for a, b in zip(first, again):
            assert np.array_equal(a["feature"][:2048], b["feature"][:2048])
            d = np.abs(a["feature"][2048:].astype(np.float64) - b["feature"][2048:].astype(np.float64)).max()
            worst = max(worst, d)
            n_diff += int(d > 0)
            n_tot += 1
print(f"first-visit vs buffered angle features: {n_diff}/{n_tot} candidates differ, max abs difference {worst:.3e}")
assert worst <= 2 ** -23                                    # one ulp of a float32 in [0.5, 1)
print("harness OK")
