"""VidToMe token merging, B200 implementation behind the reference's API
(utils/VidToMe/vidtome/__init__.py exports the same names)."""
from . import merge, patch, utils  # noqa: F401
from .patch import apply_patch, collect_from_patch, compute_merge, remove_patch, update_patch  # noqa: F401
