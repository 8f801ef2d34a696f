"""``src.utils`` as main.py uses it (utils/misc.py, utils/config.py of the reference)."""
import json
import logging
import os
import time
from collections import Counter

from ..environ import ingest
from ..utils.config import get_cfg_defaults  # noqa: F401

base_vocab = ["<PAD>", "<UNK>", "<EOS>", "<BOS>"]          # misc.py:88-89
Tokenizer = ingest.Tokenizer                                # misc.py:91-157 (encode_sentence / split_sentence / vocab_size)


def get_main_logger(log_dir=None, model_name="", save_mode="dhm"):
    """misc.py:398-435: a file logger named "main" under LOG_DIR."""
    logger = logging.getLogger("main")
    logger.setLevel(logging.INFO)
    log_dir = log_dir if log_dir is not None else os.path.join(os.path.dirname(os.getcwd()), "snapshots")
    os.makedirs(log_dir, exist_ok=True)
    fmt = {"dhm": "%Y-%m%d-%H:%M", "dh": "%Y-%m%d-%H", "d": "%Y-%m%d"}[save_mode]
    h = logging.FileHandler(os.path.join(log_dir, time.strftime(fmt, time.localtime()) + "_" + model_name + ".log"),
                            mode="a", encoding="utf-8")
    h.setLevel(logging.INFO)
    h.setFormatter(logging.Formatter(fmt="%(asctime)s - %(levelname)s: %(message)s", datefmt="%Y-%m-%d %H:%M:%S"))
    logger.addHandler(h)
    return logger


def load_datasets(splits, dataset="R2R", data_dir="tasks/R2R-judy/data"):
    """misc.py:62-69."""
    data = []
    for split in splits:
        with open("%s/%s_%s.json" % (data_dir, dataset, split)) as f:
            data += json.load(f)
    return data


def build_vocab(splits=("train",), min_count=5, start_vocab=base_vocab):
    """misc.py:188-201."""
    count = Counter()
    for item in load_datasets(list(splits)):
        for instr in item["instructions"]:
            count.update(Tokenizer.split_sentence(instr))
    vocab = list(start_vocab)
    for word, num in count.most_common():
        if num < min_count:
            break
        vocab.append(word)
    return vocab


def write_vocab(vocab, path):
    with open(path, "w") as f:
        for word in vocab:
            f.write("%s\n" % word)


def read_vocab(path):
    with open(path) as f:
        return [w.strip() for w in f.readlines()]


class FeatureSource:
    """What ``ImageFeatures.read_in`` returns here: the feature file's location plus the World (HBM tables) built from
    it on first use and shared by every environment of the run (the reference shares one in-RAM dict the same way,
    main.py:56-85)."""

    def __init__(self, path):
        self.path = path
        self.conn_dir = os.environ.get("VLN_CONNECTIVITY_DIR", "connectivity")
        self.candidates = os.environ.get("VLN_CANDIDATES_JSON", os.path.join(os.path.dirname(path) or ".", "candidates.json"))
        self._world = None

    def world(self):
        if self._world is None:
            feats = ingest.read_feature_tsv(self.path)                # decoded once: "<scan>_<viewpoint>" -> fp32 [36,2048]
            scans = sorted({long_id.split("_", 1)[0] for long_id in feats})      # = EnvBatch.featurized_scans
            world = ingest.world_from_files(self.conn_dir, scans, self.candidates, None)
            world.table = ingest.feature_table(world, feats)
            self._world = world
        return self._world


class ImageFeatures:
    """misc.py:245-312 (the part main.py touches)."""
    NUM_VIEWS, MEAN_POOLED_DIM = 36, 2048
    IMAGE_W, IMAGE_H, VFOV = 640, 480, 60

    @staticmethod
    def read_in(feature_store_path, views=36):
        return FeatureSource(feature_store_path)
