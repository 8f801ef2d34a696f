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
print("harness OK")
