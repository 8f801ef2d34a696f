"""Beam search on the GPU: the product agents' ``_dijkstra`` / ``beam_rollout`` (agent/beam.py: batched expansions on the
device-resident environment) against oracle/port_beam.py, which tests/_ref_check_beam.py pins to the unmodified
reference's ``_dijkstra``.  Same K best paths per episode (viewpoints, poses, actions), listener scores within 1e-3,
the same navigation path; speaker rescoring of those paths within 1e-3."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def _key(p):
    return (tuple(p["action"]), tuple(x[0] for x in p["trajectory"]))


@pytest.mark.parametrize("kind", ["ENVDROP", "FOLLOWER"])
def test_dijkstra_matches_oracle(kind):
    from oracle import port_beam as PB
    from test_agents_gpu import _setup
    agent, pag, env, penv, sds, cfg = _setup(kind, B=6)
    agent.eval()
    for K in (3, 5):
        with torch.no_grad():
            got = agent._dijkstra(K)
            ref = PB.dijkstra(pag, penv, K)
        assert [r["instr_id"] for r in got] == [r["instr_id"] for r in ref]
        same = total = 0
        for g, r in zip(got, ref):
            assert g["scan"] == r["scan"] and list(g["instr_encoding"]) == list(r["instr_encoding"])
            gp, rp = {_key(p): p for p in g["paths"]}, {_key(p): p for p in r["paths"]}
            assert len(g["paths"]) == len(r["paths"]) == len(gp)
            total += len(rp)
            for k, a in gp.items():
                if k not in rp:                       # (a near-tie between the K-th and (K+1)-th path may swap them)
                    continue
                b = rp[k]
                same += 1
                assert a["trajectory"] == b["trajectory"]
                assert a["listener_actions"] == b["listener_actions"]
                assert np.allclose(a["listener_scores"], b["listener_scores"], rtol=1e-3, atol=1e-4)
                assert len(a["visual_feature"]) == len(b["visual_feature"])
            if set(gp) == set(rp):
                assert g["dijk_path"] == r["dijk_path"]
        assert same >= 0.95 * total, (same, total)


def test_beam_rollout_speaker_scores_match_oracle():
    """beam_rollout: the listener's K best paths rescored by the speaker (dropout off so that the reference's
    train-mode call is deterministic) == the oracle's rescoring of the same paths."""
    from clvln_b200 import utils
    from clvln_b200.agent import Speaker
    from oracle import port_beam as PB, port_speaker as PS
    from test_agents_gpu import _setup
    agent, pag, env, penv, sds, cfg = _setup("ENVDROP", B=4)
    agent.eval()
    scfg = utils.get_cfg_defaults().AIDE.SPEAKER
    scfg.DROPOUT, scfg.FEAT_DROPOUT = 0.0, 0.0
    torch.manual_seed(3)
    spk = Speaker(scfg, agent.device, agent.tokenizer, env=env)
    port = PS.SpeakerPort({k: v.detach().cpu() for k, v in spk.encoder.state_dict().items()},
                          {k: v.detach().cpu() for k, v in spk.decoder.state_dict().items()}, p=0.0, pf=0.0)
    with torch.no_grad():
        got = agent.beam_rollout(spk, 3)
        ref = PB.dijkstra(pag, penv, 3)
    checked = 0
    for g, r in zip(got, ref):
        assert g["instr_id"] == r["instr_id"]
        with torch.no_grad():
            want = PB.speaker_scores(port, r)
        rp = {_key(p): s for p, s in zip(r["paths"], want)}
        for p in g["paths"]:
            assert "visual_feature" not in p
            if _key(p) in rp:
                s = rp[_key(p)]
                assert p["speaker_scores"].shape == s.shape
                assert np.allclose(p["speaker_scores"], s, rtol=1e-3, atol=1e-3)
                checked += 1
    assert checked >= 10
    env.reset_epoch()
    agent.beam_search(spk, beam_size=2)
    assert len(agent.results) == len(env.data) and all("paths" in v for v in agent.results.values())
