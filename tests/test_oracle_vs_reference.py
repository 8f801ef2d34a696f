"""Pin oracle/port_*.py against the REAL reference (container only; skipped where
/root/reference is absent, e.g. on the GPU box — there tests/golden/ carries the pin)."""
import os
import runpy
import sys

import pytest

from oracle import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = [pytest.mark.ref,
              pytest.mark.skipif(not ref_loader.reference_available(), reason="no /root/reference")]


def test_port_modules_match_reference_modules():
    """EncoderLSTM (3 configs), EnvDrop/Follower/Monitor decoders, Critic: eval + train(RNG-matched)."""
    runpy.run_path(os.path.join(HERE, "_ref_check_modules.py"), run_name="__main__")


def test_reference_env_matches_world_tables():
    """The reference's own R2RBatch/make_candidate over FakeSim == World index tables, bit-exact."""
    runpy.run_path(os.path.join(HERE, "_ref_check_env.py"), run_name="__main__")


def test_port_rollouts_match_reference_agents():
    """Real Follower/Monitor/EnvDrop agents' rollout + backward == port: loss, grads, trajectories."""
    runpy.run_path(os.path.join(HERE, "_ref_check_rollout.py"), run_name="__main__")


def test_port_speaker_matches_reference_speaker():
    """Real SpeakerEncoder / SpeakerDecoder (eval, RNG-matched train, gradients) and the real Speaker class
    (from_shortest_path, teacher_forcing incl. the beam-search entry point, greedy infer_batch) == oracle/port_speaker.py."""
    runpy.run_path(os.path.join(HERE, "_ref_check_speaker.py"), run_name="__main__")


def test_port_beam_search_matches_reference_dijkstra():
    """Real EnvDrop / Follower agents' _dijkstra (K best listener paths, base.py:183-397) == oracle/port_beam.py: paths,
    actions, scores, visual features, dijk_path."""
    runpy.run_path(os.path.join(HERE, "_ref_check_beam.py"), run_name="__main__")


def test_ingest_matches_reference_loaders():
    """environ/ingest.py == ImageFeatures.read_in, load_nav_graphs + networkx paths (ties), Tokenizer."""
    runpy.run_path(os.path.join(HERE, "_ref_check_ingest.py"), run_name="__main__")


def test_evaluator_matches_reference_evaluation():
    """engine/evaluator.py (nav/oracle error, SPL, nDTW, SDTW, CLS, rates) == src/engine/evaluator.py Evaluation.score."""
    runpy.run_path(os.path.join(HERE, "_ref_check_eval.py"), run_name="__main__")


def test_round_sizes_are_the_shipped_clr2r_rounds():
    """split_rounds' default proportions (synthetic stand-in for the offline CLR2R difficulty split) are the instruction
    counts of the reference's data/CLR2R/CLR2R_train_round[k]_v3.json files: 1037 / 1415 / 4897 / 4593 / 2097 = 14 039."""
    import inspect
    import json
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import split_rounds
    sizes = inspect.signature(split_rounds).parameters["sizes"].default
    data_dir = os.path.join(ref_loader.REF_TASK, "data", "CLR2R")
    counted = []
    for k in range(1, 6):
        with open(os.path.join(data_dir, f"CLR2R_train_round[{k}]_v3.json")) as f:
            counted.append(sum(len(it["instructions"]) for it in json.load(f)))
    assert tuple(counted) == tuple(sizes) and sum(counted) == 14039


def test_reference_yaml_configs_load_into_the_config_tree():
    """Every configs/*/*.yaml of the reference merges into utils.get_cfg_defaults() (same keys as src/utils/config.py),
    and the model node the agents read is there for each."""
    import glob
    import clvln_b200  # noqa: F401
    from clvln_b200 import utils
    files = sorted(glob.glob(os.path.join(ref_loader.REF_TASK, "configs", "*", "*.yaml")))
    assert len(files) >= 6
    seen = set()
    for f in files:
        cfg = utils.get_cfg_defaults()
        cfg.merge_from_file(f)
        assert cfg.MODEL.NAME in ("ENVDROP", "FOLLOWER", "SELF-MONITOR") and cfg.TRAIN.CLMODE in ("", "NAIVE", "SELF-PACE")
        node = {"ENVDROP": "ENVDROP", "FOLLOWER": "FOLLOWER", "SELF-MONITOR": "MONITOR"}[cfg.MODEL.NAME]
        mc = getattr(cfg.MODEL, node)
        assert mc.HIDDEN_SIZE in (256, 512) and cfg.TRAIN.BATCH_SIZE > 0 and cfg.AGENT.MAX_EPISODE_LEN > 0
        seen.add((cfg.MODEL.NAME, cfg.TRAIN.CLMODE))
    assert {n for n, _ in seen} == {"ENVDROP", "FOLLOWER", "SELF-MONITOR"}


def test_state_dicts_are_interchangeable_with_the_reference_modules():
    """Checkpoint compatibility: each of the five modules has exactly the reference module's state_dict keys
    and shapes, so `load_state_dict(strict=True)` works in both directions (EnvDrop / Follower / Self-Monitor shapes)."""
    import torch
    import clvln_b200  # noqa: F401
    from clvln_b200 import model as M
    RU, RP = ref_loader.load_ref_models()
    pairs = [
        (M.EncoderLSTM(992, 256, 512, 0, 0.5, True, 1), RU.EncoderLSTM(992, 256, 512, 0, 0.5, True, 1)),      # EnvDrop
        (M.EncoderLSTM(992, 300, 256, 0, 0.5, True, 2), RU.EncoderLSTM(992, 300, 256, 0, 0.5, True, 2)),      # Follower
        (M.EncoderLSTM(992, 256, 512, 0, 0.5, False, 1), RU.EncoderLSTM(992, 256, 512, 0, 0.5, False, 1)),    # Self-Monitor
        (M.EnvDropDecoder(512, 0.5, 0.3, 64), RP.EnvDropDecoder(512, 0.5, 0.3, 64)),
        (M.AttnDecoderLSTM(256, 0.5), RP.AttnDecoderLSTM(256, 0.5)),
        (M.MonitorDecoder(512, 0.5, 80, [1024]), RP.MonitorDecoder(512, 0.5, 80, [1024])),
        (M.Critic(512, 0.5), RP.Critic(512, 0.5)),
    ]
    for mine, ref in pairs:
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys()), (type(mine).__name__, sorted(set(a) ^ set(b)))
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
        mine.load_state_dict(b, strict=True)
        ref.load_state_dict(mine.state_dict(), strict=True)
        assert all(torch.equal(mine.state_dict()[k], ref.state_dict()[k]) for k in a)


def test_naive_curriculum_schedule_matches_reference():
    """NaiveCurriculum.curriculum_strategy (curriculum.py:176-179): same round per epoch as the real class."""
    import clvln_b200  # noqa: F401
    from clvln_b200.engine import NaiveCurriculum
    ref_loader.load_ref_agents()
    from src.engine.curriculum import NaiveCurriculum as RefNaive
    envs = {f"round_{k}": k for k in range(1, 6)}
    for sw in (1, 7, 20):
        ref, mine = RefNaive(switch_epoch=sw), NaiveCurriculum(switch_epoch=sw)
        for ep in range(1, 8 * sw + 3):
            assert mine.curriculum_strategy(envs, ep) == ref.curriculum_strategy(envs, ep)
