"""VidToMe token merging, B200 implementation behind the reference's API
(utils/VidToMe/vidtome/__init__.py exports the same names)."""
from . import merge, patch, utils  # noqa: F401
from .patch import (apply_patch, assert_draws_consumed, collect_from_patch, compute_merge, prefetch_draws,  # noqa: F401
                    remove_patch, update_patch)
