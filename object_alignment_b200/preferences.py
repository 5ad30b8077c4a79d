"""The ICP knobs of ObjectAlignmentPreferences (reference lib/preferences.py:27-72), same names,
defaults and ranges.  Inside Blender `get_addon_preferences()` returns the addon's real
preferences object; elsewhere it returns the module-level `IcpPreferences` instance that
`set_addon_preferences()` installed (defaults otherwise)."""
from __future__ import annotations

try:
    import bpy  # type: ignore
except ImportError:
    bpy = None


class IcpPreferences:
    """Plain-Python mirror of the nine ICP properties (lib/preferences.py:31-72)."""

    def __init__(self, **kw):
        self.icp_iterations = 50        # IntProperty default 50          (:31-33)
        self.redraw_frequency = 10      # modal operator only             (:34-37)
        self.use_sample = False         # defined but never read          (:38-41)
        self.sample_fraction = 0.5      # [0, 1] -> stride round(1/x)     (:42-47)
        self.min_start = 0.5            # [0, 20] world-distance threshold(:48-53)
        self.target_d = 0.01            # [0, 10] translation target      (:54-59)
        self.use_target = True          #                                 (:60-63)
        self.take_m_with = False        #                                 (:64-67)
        self.align_meth = '0'           # '0' RIGID, '1' ROT_LOC_SCALE    (:68-72)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("unknown ICP preference %r" % k)
            setattr(self, k, v)


_prefs = None


def set_addon_preferences(prefs):
    global _prefs
    _prefs = prefs


def get_addon_preferences():
    """reference functions/common/blender.py:48-55"""
    global _prefs
    if _prefs is not None:
        return _prefs
    if bpy is not None:
        addon = bpy.context.preferences.addons.get(__package__.split(".")[0])
        if addon is not None and addon.preferences:
            return addon.preferences
    _prefs = IcpPreferences()
    return _prefs
