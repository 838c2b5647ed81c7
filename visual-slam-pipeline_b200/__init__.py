"""B200-native descriptor matching engine: host-side mirror of the reference's matcher
call sites (Slam::match_features, LoopCloser::detect's matching block, the map-point DB
search) over the C ABI of libvsm.so.  See include/vsm.h and DESIGN.md.

The directory name carries a hyphen, so import it through the repo-root shim:
    import vsm_b200
"""
from .matcher import (DMATCH, Group, Matcher, TrackCfg, VsmError, lib_path, load_library,  # noqa: F401
                      ENGINE_AUTO, ENGINE_TENSOR, ENGINE_SIMT, ENGINE_TENSOR_PAIR)


def load_sharded():
    """The multi-GPU database search (imports torch lazily)."""
    from . import sharded as _s
    return _s
