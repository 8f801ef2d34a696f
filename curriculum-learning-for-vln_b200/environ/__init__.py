from . import world  # noqa: F401
from .world import World, make_world, make_items  # noqa: F401
