from . import world  # noqa: F401
from .world import World, make_world, make_items, full_world_sizes  # noqa: F401
from .batch import R2RBatch, CLR2RBatch, IndexBatch, split_rounds  # noqa: F401
