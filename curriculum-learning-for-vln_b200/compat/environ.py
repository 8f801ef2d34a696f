"""``src.environ`` with the reference's constructor signatures (common_env.py:120, curriculum_env.py:30), built from
the reference's on-disk data through environ/ingest.py."""
import torch

from ..environ import batch as _batch
from ..environ import ingest
from . import utils as _utils


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None


def _dp():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class R2RBatch(_batch.R2RBatch):
    def __init__(self, feature_store, batch_size=100, splits=("train",), tokenizer=None, data_name="R2R",
                 data_dir="tasks/R2R-judy/data"):
        world = feature_store.world()
        items = []
        for split in splits:
            items += ingest.items_from_r2r_json("%s/%s_%s.json" % (data_dir, data_name, split), world, tokenizer)
        rank, ws = _dp()
        super().__init__(world, items, batch_size=batch_size, name=splits[0] if len(splits) else "FAKE", device=_device(),
                         max_len=getattr(tokenizer, "encoding_length", 80), rank=rank, world_size=ws)
        self.splits = list(splits)
        self.tok = tokenizer


class CLR2RBatch(_batch.CLR2RBatch):
    def __init__(self, feature_store, batch_size=100, c_rate=0.8, tokenizer=None, data_dir="tasks/R2R-judy/data/CLR2Rv3"):
        world = feature_store.world()
        rounds = {k: ingest.items_from_r2r_json("%s/CLR2R_train_round[%d]_v3.json" % (data_dir, k), world, tokenizer)
                  for k in range(1, 6)}
        rank, ws = _dp()
        super().__init__(world, rounds, batch_size=batch_size, c_rate=c_rate, device=_device(),
                         max_len=getattr(tokenizer, "encoding_length", 80), rank=rank, world_size=ws)
        self.tok = tokenizer


load_datasets = _utils.load_datasets
