from . import world  # noqa: F401
from .world import World, make_world, make_items, full_world_sizes  # noqa: F401
from .batch import R2RBatch, CLR2RBatch, IndexBatch, split_rounds  # noqa: F401
from . import ingest  # noqa: F401,E402  (real-data conversion: TSV / connectivity / R2R json -> World tables)
