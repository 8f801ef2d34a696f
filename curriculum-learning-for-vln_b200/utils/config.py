"""Config tree with the reference's keys and defaults (src/utils/config.py:3-126) on a small
yacs-compatible CfgNode (yacs is not installed here), plus two helpers of src/utils/misc.py the
hot path touches: the tokenizer surface the agents use and length2mask (:481-486)."""
import copy

import torch


class CfgNode(dict):
    """Attribute-access dict with merge_from_file / merge_from_list / clone / freeze."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        return self

    @staticmethod
    def _decode(v):
        """yacs' _decode_cfg_value: strings that are Python literals become them ("(1024, )" -> (1024,), as
        configs/monitor/*.yaml write MLP_HIDDEN); anything else stays a string."""
        if isinstance(v, str):
            import ast
            try:
                return ast.literal_eval(v)
            except Exception:
                return v
        return v

    def merge_from_other(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_from_other(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else self._decode(v)

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self.merge_from_other(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        import ast
        assert len(lst) % 2 == 0
        for key, val in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = self._decode(val)


def get_cfg_defaults():
    C = CfgNode
    cfg = C()
    cfg.DATA = C(dict(NAME="R2R", DATA_DIR="tasks/R2R-judy/data", TRAIN_VOCAB="", TRAINVAL_VOCAB="", MAX_ENC_LEN=20,
                      MAX_SUBINSTR_NUM=0, IMG_FEAT_DIR=""))
    cfg.TRAIN = C(dict(DEVICE=0, OPTIM="", LR=1e-4, BATCH_SIZE=128, START_EPOCH=1, MAX_EPOCH=0, ITER_PER_EPOCH=200,
                       EVAL_INTERVAL=1, SCHEDULER="", PATIENCE=3, LR_MIN=1e-6, DATA_ARGUMENT=False,
                       PROGMONITOR_WEIGHT=0.5, EVAL_TRAIN=False, CLMODE="",
                       SELF_PACE=dict(CRATE=1.0, WCTRL=0.0, LAMB=0.0, MIU=0.0, FUNC="", BURN_IN=0, INTERVAL=0,
                                      STRATEGY=""),
                       AUTO_CULM=dict(ALPHA=0.0, ETA=0.0, BETA=0.0, EPS=0.0, RRSIZE=0)))
    cfg.OUTPUT = C(dict(RESUME="", CKPT_DIR="", LOG_DIR="", RESULT_DIR="", TSBOARD=1, TSBOARD_DIR=""))
    cfg.AGENT = C(dict(TEACHER_FORCE=False, MAX_EPISODE_LEN=20, FEEDBACK="sample"))
    cfg.MODEL = C(dict(
        NAME="",
        FOLLOWER=dict(GLOVE_PATH="", WORD_EMB_SIZE=0, HIDDEN_SIZE=0, DROP_RATE=0.5, ENC_BIDIRECTION=True,
                      ENC_LAYERS=1),
        MONITOR=dict(WORD_EMB_SIZE=0, HIDDEN_SIZE=0, DROP_RATE=0.5, ENC_BIDIRECTION=True, ENC_LAYERS=1,
                     MLP_HIDDEN=(128,)),
        ENVDROP=dict(WORD_EMB_SIZE=0, ACT_EMB_SIZE=0, HIDDEN_SIZE=0, DROP_RATE=0.5, FEAT_DROP_RATE=0.3,
                     ENC_BIDIRECTION=True, ENC_LAYERS=1, ML_WEIGHT=0.0, GAMMA=0.0, RL_NORMALIZE="none")))
    cfg.AIDE = C(dict(SPEAKER=dict(RNN_DIM=512, DROPOUT=0.6, FEAT_DROPOUT=0.3, BI_DIRECTION=True, WEMB=256, LR=1e-4,
                                   FAST_TRAIN=False, IGNORE_ID=-1, MAX_DECODE=120, LOAD_OPTIM=False)))    # config.py:109-119
    return cfg


def agent_cfg(name, **over):
    """The shipped configs (configs/{follower,monitor,envdrop}/*.yaml) as a full config tree."""
    cfg = get_cfg_defaults()
    cfg.DATA.MAX_ENC_LEN = 80
    cfg.TRAIN.BATCH_SIZE = 64
    cfg.TRAIN.LR = 1e-4
    if name == "ENVDROP":
        cfg.MODEL.NAME = "ENVDROP"
        cfg.MODEL.ENVDROP.merge_from_other(dict(WORD_EMB_SIZE=256, ACT_EMB_SIZE=64, HIDDEN_SIZE=512, DROP_RATE=0.5,
                                                FEAT_DROP_RATE=0.3, ENC_BIDIRECTION=True, ENC_LAYERS=1,
                                                ML_WEIGHT=0.2, GAMMA=0.9, RL_NORMALIZE="total"))
        cfg.TRAIN.OPTIM, cfg.AGENT.MAX_EPISODE_LEN = "rms", 35      # configs/envdrop/envdrop_config.yaml:18
    elif name == "FOLLOWER":
        cfg.MODEL.NAME = "FOLLOWER"
        cfg.MODEL.FOLLOWER.merge_from_other(dict(WORD_EMB_SIZE=300, HIDDEN_SIZE=256, DROP_RATE=0.5,
                                                 ENC_BIDIRECTION=True, ENC_LAYERS=2))
        cfg.TRAIN.OPTIM, cfg.AGENT.MAX_EPISODE_LEN = "adam", 10
    elif name in ("SELF-MONITOR", "MONITOR"):
        cfg.MODEL.NAME = "SELF-MONITOR"
        cfg.MODEL.MONITOR.merge_from_other(dict(WORD_EMB_SIZE=256, HIDDEN_SIZE=512, DROP_RATE=0.5,
                                                ENC_BIDIRECTION=False, ENC_LAYERS=1, MLP_HIDDEN=[1024]))
        cfg.TRAIN.OPTIM, cfg.AGENT.MAX_EPISODE_LEN = "adam", 10
    else:
        raise KeyError(name)
    if over:
        cfg.merge_from_list([x for kv in over.items() for x in kv])
    return cfg


class StubTokenizer:
    """The Tokenizer surface the agents touch (misc.py:94-184): vocab_size() and word_to_index.
    Synthetic items already carry their token ids."""

    def __init__(self, vocab=992):
        self._vocab = vocab
        self.word_to_index = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<BOS>": 3}

    def vocab_size(self):
        return self._vocab


def length2mask(length, device=None, size=None):
    """misc.py:481-486: mask[i,j] = j >= length[i]."""
    length = torch.as_tensor(length)
    size = int(length.max()) if size is None else size
    mask = torch.arange(size, dtype=torch.int64).unsqueeze(0) >= length.long().cpu().unsqueeze(1)
    return mask.to(device) if device is not None else mask
