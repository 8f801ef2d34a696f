"""Drop-in at the level the north star names: "main.py uses it as a drop-in".

clvln_b200.compat offers the reference's `src` package interface (utils / environ / engine / agent with the reference's
constructor signatures, built from the reference's own on-disk formats).  Here
  * (container, CPU) the reference's UNMODIFIED tasks/R2R-judy/main.py is executed with `src` bound to the facade: vocab /
    tokenizer / ImageFeatures.read_in / R2RBatch + CLR2RBatch construction / build_agent / trainer selection /
    `trainer.train(cfg, agent, cfg.OUTPUT.TSBOARD_DIR, train_env, valid_env)` all run (MAX_EPOCH = 0: no GPU here);
  * (GPU box, no reference needed) the same sequence as main.py:50-125, written out against the facade, trains for an
    epoch with validation, scoring and checkpoints.
"""
import io
import os
import runpy
import sys
from contextlib import redirect_stdout

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAIN = "/root/reference/tasks/R2R-judy/main.py"
REF_CFG = "/root/reference/tasks/R2R-judy/configs"


def _dataset(tmp_path, clr2r=False):
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import ingest, make_world, make_items, split_rounds
    world = make_world(n_scans=2, seed=4)
    items = make_items(world, 72, seed=4)
    splits = {"train": items[:48], "val_seen": items[48:60], "val_unseen": items[60:]}
    paths = ingest.write_reference_dataset(world, splits, str(tmp_path))
    if clr2r:
        rounds = split_rounds(items[:48])
        ingest.write_reference_dataset(world, {"train_round[%d]_v3" % k: rounds[k] for k in range(1, 6)}, str(tmp_path),
                                       dataset="CLR2R", data_dir="tasks/R2R-judy/data/CLR2Rv3")
    return world, splits, paths


def _overrides(tmp_path, paths):
    return ["DATA.IMG_FEAT_DIR", paths["tsv"], "DATA.TRAIN_VOCAB", paths["train_vocab.txt"],
            "DATA.TRAINVAL_VOCAB", paths["trainval_vocab.txt"], "OUTPUT.LOG_DIR", str(tmp_path / "logs"),
            "OUTPUT.CKPT_DIR", str(tmp_path / "ckpt"), "OUTPUT.RESULT_DIR", str(tmp_path / "trajs"),
            "OUTPUT.TSBOARD_DIR", str(tmp_path / "tb"), "TRAIN.BATCH_SIZE", "8"]


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="needs the reference's main.py (build container only)")
@pytest.mark.parametrize("yaml_name,clmode", [("envdrop/envdrop_config.yaml", ""), ("envdrop/envdrop_cl_config.yaml", "NAIVE"),
                                              ("envdrop/envdrop_cl_config.yaml", "SELF-PACE"),
                                              ("follower/follower_config.yaml", ""), ("monitor/selfmonitor_config.yaml", "")])
def test_reference_main_py_runs_unchanged_on_the_facade(tmp_path, monkeypatch, yaml_name, clmode):
    import clvln_b200.compat as compat
    from clvln_b200 import engine as E, agent as A
    world, splits, paths = _dataset(tmp_path, clr2r=bool(clmode))
    monkeypatch.chdir(tmp_path)                       # main.py's relative paths: connectivity/, tasks/R2R-judy/data/...
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    calls = []
    orig_train = E.ClassicTrainer.train

    def train(self, cfg, agent, tsboard_dir, train_env, valid_env, *a, **kw):
        calls.append((self, cfg, agent, tsboard_dir, train_env, valid_env))
        return orig_train(self, cfg, agent, tsboard_dir, train_env, valid_env, *a, **kw)
    monkeypatch.setattr(E.ClassicTrainer, "train", train)
    argv = ["main.py", "--config-file", os.path.join(REF_CFG, yaml_name), "--seed", "2020"] + _overrides(tmp_path, paths) + \
        ["TRAIN.MAX_EPOCH", "0", "TRAIN.CLMODE", clmode or "''"]
    monkeypatch.setattr(sys, "argv", argv)
    out = io.StringIO()
    try:
        compat.install_as_src()
        with redirect_stdout(out):
            runpy.run_path(REF_MAIN, run_name="__main__")
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    text = out.getvalue()
    assert "Traceback" not in text, text[-3000:]
    assert "[5] Training Finished!" in text and len(calls) == 1
    trainer, cfg, agent, tsb, train_env, valid_env = calls[0]
    kind = {"": E.ClassicTrainer, "NAIVE": E.NaiveCurriculum, "SELF-PACE": E.SelfPacedCurriculum}[clmode]
    assert type(trainer) is kind
    assert isinstance(agent, {"ENVDROP": A.EnvDropAgent, "FOLLOWER": A.FollowerAgent, "SELF-MONITOR": A.SelfMonitorAgent}[cfg.MODEL.NAME])
    assert tsb == str(tmp_path / "tb") and set(valid_env) == {"val_seen", "val_unseen"}
    assert [len(e.data) for e in valid_env.values()] == [12, 12]
    if clmode == "NAIVE":
        assert set(train_env) == {f"round_{k}" for k in range(1, 6)} and len(train_env["round_5"].data) == 48
    else:
        assert len(train_env.data) == 48
        it = train_env.data[0]
        assert it["instr_encoding"][0] == 3 and it["instr_length"] >= 5                   # tokenised with the vocab on disk
    if clmode == "SELF-PACE":
        assert len(trainer.weight) == 48 and float(trainer.c) == float(train_env.a.sum())   # CRATE 1.0 in the shipped yaml
    assert cfg.TRAIN.OPTIM in ("rms", "adam")


@pytest.mark.gpu
def test_main_py_sequence_trains_evaluates_and_checkpoints(tmp_path, monkeypatch):
    """main.py:50-125 written out against the facade (the GPU box has no reference): tokenizer from the vocab file,
    ImageFeatures.read_in, R2RBatch(feature_store, batch_size, splits=, tokenizer=), build_agent, ClassicTrainer().train(cfg,
    agent, TSBOARD_DIR, train_env, valid_env) — one epoch with validation (the default evaluator), best / latest checkpoints."""
    import random
    import numpy as np
    import clvln_b200.compat as compat
    world, splits, paths = _dataset(tmp_path)
    monkeypatch.chdir(tmp_path)
    src = compat.install_as_src()
    utils, engine, environ = src.utils, src.engine, src.environ
    from clvln_b200.utils import agent_cfg
    cfg = agent_cfg("ENVDROP")              # get_cfg_defaults() + the values of configs/envdrop/envdrop_config.yaml (not on this box)
    cfg.merge_from_list(_overrides(tmp_path, paths) + ["MODEL.NAME", "ENVDROP", "TRAIN.OPTIM", "rms", "TRAIN.MAX_EPOCH", "1",
                                                       "TRAIN.ITER_PER_EPOCH", "3", "TRAIN.EVAL_INTERVAL", "1",
                                                       "AGENT.MAX_EPISODE_LEN", "10", "AGENT.FEEDBACK", "sample"])
    random.seed(2020), np.random.seed(2020), torch.manual_seed(2020)
    device = torch.device("cuda:%d" % cfg.TRAIN.DEVICE)
    tok = utils.Tokenizer(utils.read_vocab(cfg.DATA.TRAIN_VOCAB), cfg.DATA.MAX_ENC_LEN)
    img_feature = utils.ImageFeatures.read_in(cfg.DATA.IMG_FEAT_DIR)
    train_env = environ.R2RBatch(img_feature, cfg.TRAIN.BATCH_SIZE, splits=["train"], tokenizer=tok)
    valid_env = {k: environ.R2RBatch(img_feature, cfg.TRAIN.BATCH_SIZE, splits=[k], tokenizer=tok) for k in ("val_seen", "val_unseen")}
    assert train_env.world is valid_env["val_seen"].world                   # one feature table for every environment
    agent = src.agent.build_agent(cfg, tok, device)
    trainer = engine.ClassicTrainer()
    trainer.train(cfg, agent, cfg.OUTPUT.TSBOARD_DIR, train_env, valid_env)
    h = trainer.history
    assert len(h) == 1 and np.isfinite(h[0]["loss_avg"])
    for k in ("val_seen", "val_unseen"):
        assert 0.0 <= h[0][k]["success_rate"] <= 1.0 and np.isfinite(h[0][k]["spl"])
    files = os.listdir(tmp_path / "ckpt")
    assert any(f.startswith("latest_avgloss:") for f in files)
