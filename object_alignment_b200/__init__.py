"""object_alignment_b200 -- B200-native ICP alignment engine behind the operator surface of
patmo141/object_alignment (operators/icp_align.py, functions/general.py).

Python here is glue only: ctypes -> libicp_b200.so (hand-written sm_100a CUDA, optional NCCL).
Importing the package does not need a GPU; calling into it does, and fails loudly without one.
"""
from .engine import IcpEngine, IcpError, PinnedArray, shard_range, stride_from_sample_fraction
from .functions import (SurfaceNN, VertexNN, affine_matrix_from_points, base_index_from_object,
                        calc_target, icp_step, make_pairs)
from .operators import (OBJECT_OT_align_pick_points, OBJECT_OT_icp_align, OBJECT_OT_icp_align_feedback,
                        build_vlist, build_vlist_from_arrays)
from .preferences import IcpPreferences, get_addon_preferences, set_addon_preferences

__all__ = [
    "IcpEngine", "IcpError", "PinnedArray", "shard_range", "stride_from_sample_fraction",
    "VertexNN", "SurfaceNN", "base_index_from_object", "affine_matrix_from_points", "calc_target", "icp_step", "make_pairs",
    "OBJECT_OT_icp_align", "OBJECT_OT_icp_align_feedback", "OBJECT_OT_align_pick_points", "build_vlist",
    "build_vlist_from_arrays",
    "IcpPreferences", "get_addon_preferences", "set_addon_preferences",
]
__version__ = "0.1.0"
