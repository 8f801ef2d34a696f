"""`src`-compatible facade: the reference's package interface as its own ``main.py`` uses it
(tasks/R2R-judy/main.py:10-12 ``from src import utils, engine, environ``, ``from src.utils import ImageFeatures``,
``from src.agent import build_agent``), implemented on this repository's device-resident path.

    import clvln_b200.compat as compat
    compat.install_as_src()          # sys.modules["src"], "src.utils", "src.engine", "src.environ", "src.agent"
    runpy.run_path("tasks/R2R-judy/main.py", run_name="__main__")        # the reference's CLI, unchanged

What each name maps to (same call shapes as the reference, main.py:33-133):
  utils.get_cfg_defaults / get_main_logger / read_vocab / write_vocab / build_vocab / Tokenizer / ImageFeatures.read_in
  environ.R2RBatch(feature_store, batch_size, splits=, tokenizer=, data_name=, data_dir=)      common_env.py:120
  environ.CLR2RBatch(feature_store, batch_size, c_rate=, tokenizer=, data_dir=)                 curriculum_env.py:30
  engine.ClassicTrainer() / NaiveCurriculum() / SelfPacedCurriculum(train_env, device, pace_func=, ...) / check_the_code
  agent.build_agent(cfg, tok, device)
The environments are built from the reference's own on-disk formats through environ/ingest.py: the ResNet TSV
(``ImageFeatures.read_in`` keeps the path; nothing is decoded until the first env needs the table), connectivity/*.json,
R2R / CLR2R json, and — because the candidate geometry is defined only inside the Matterport simulator — a JSON dump of
the reference's candidate cache (``ingest.dump_candidates_with_reference``), looked up next to the TSV
(``<tsv dir>/candidates.json``) or at $VLN_CANDIDATES_JSON; connectivity files at ./connectivity or $VLN_CONNECTIVITY_DIR.
"""
import sys
import types

from . import agent, engine, environ, utils  # noqa: F401


def install_as_src():
    """Register this facade under the reference's package name."""
    pkg = sys.modules[__name__]
    sys.modules["src"] = pkg
    for name in ("utils", "engine", "environ", "agent"):
        sys.modules[f"src.{name}"] = getattr(pkg, name)
    model = types.ModuleType("src.model")
    from .. import model as _model
    for k in ("EncoderLSTM", "AttnDecoderLSTM", "MonitorDecoder", "EnvDropDecoder", "Critic", "SpeakerEncoder", "SpeakerDecoder"):
        setattr(model, k, getattr(_model, k))
    sys.modules["src.model"] = model
    pkg.model = model
    return pkg
