"""OBJECT_OT_icp_align -- the reference's operator surface over the GPU engine.

Same class attributes and methods as reference operators/icp_align.py:32-161
(``bl_idname``/``bl_label``/``bl_options``, ``poll``, ``execute``); registrable by the same
``bpy.utils.register_class`` loop (reference __init__.py:43-53).  Outside Blender (tests, the
GPU box) ``bpy`` is absent and ``Operator`` is a plain base class.

execute() keeps the reference's observable behaviour -- same vertex-group rule, same
parameters, same per-iteration pose update and convergence test -- but runs the whole
``while n < iters and not converged`` loop (icp_align.py:96-151) on the GPU through
``icp_run`` and writes ``matrix_world`` back once.  The depsgraph refresh the reference does
every iteration (icp_align.py:132-133) happens once at the end.
"""
from __future__ import annotations

import time

import numpy as np

from . import _native as N
from .engine import IcpError, PinnedArray, stride_from_sample_fraction
from .functions import (affine_matrix_from_points, array_to_matrix_like, base_index_from_object,
                        extract_coords, matrix_to_array)
from .preferences import get_addon_preferences

try:  # inside Blender
    import bpy  # type: ignore
    from bpy.types import Operator  # type: ignore
except ImportError:  # anywhere else
    bpy = None

    class Operator:  # minimal stand-in so the class body below is importable
        def report(self, level, message):
            print("%s: %s" % ("/".join(sorted(level)), message))


def build_vlist(align_obj) -> np.ndarray:
    """Vertex-group rule of the reference (operators/icp_align.py:56-80) -> ascending int32.

    'icp_include' wins over 'icp_exclude' (the ``elif`` at :67); include keeps members with
    weight > 0.9 (:64); exclude keeps non-members (:71) and members with weight < 0.1 (:75).
    """
    verts = align_obj.data.vertices
    lookup = {g.name: g.index for g in align_obj.vertex_groups}
    n = len(verts)
    if "icp_include" in align_obj.vertex_groups:
        gi = lookup["icp_include"]
        keep = np.zeros(n, dtype=bool)
        for v in verts:
            for g in v.groups:
                if g.group == gi and g.weight > 0.9:
                    keep[v.index] = True
        return np.nonzero(keep)[0].astype(np.int32)
    if "icp_exclude" in align_obj.vertex_groups:
        gi = lookup["icp_exclude"]
        keep = np.zeros(n, dtype=bool)
        for v in verts:
            member = False
            for g in v.groups:
                if g.group == gi:
                    member = True
                    if g.weight < 0.1:
                        keep[v.index] = True
            if not member:
                keep[v.index] = True
        return np.nonzero(keep)[0].astype(np.int32)
    return np.arange(n, dtype=np.int32)


def build_vlist_from_arrays(n_verts, groups) -> np.ndarray:
    """Same rule on bulk arrays: groups = {name: (member_mask bool[n], weight float32[n])}."""
    groups = groups or {}
    if "icp_include" in groups:
        mask, w = groups["icp_include"]
        keep = np.asarray(mask, bool) & (np.asarray(w, np.float32).astype(np.float64) > 0.9)
    elif "icp_exclude" in groups:
        mask, w = groups["icp_exclude"]
        mask = np.asarray(mask, bool)
        keep = (~mask) | (np.asarray(w, np.float32).astype(np.float64) < 0.1)
    else:
        keep = np.ones(int(n_verts), dtype=bool)
    return np.nonzero(keep)[0].astype(np.int32)


def _apply_to_m_objects(context, new_mats):
    """take_m_with (icp_align.py:123-127): post-multiply every scene object named 'm_*'."""
    scene_objects = None
    if getattr(context, "scene", None) is not None:
        scene_objects = context.scene.objects
    elif bpy is not None:
        scene_objects = bpy.context.scene.objects        # what the reference reads (icp_align.py:124)
    if not scene_objects:
        return
    for obj in scene_objects:
        if obj.name[:2] == "m_":
            for nm in new_mats:
                obj.matrix_world = obj.matrix_world @ array_to_matrix_like(nm, obj.matrix_world)
            obj.update_tag()


def upload_moving_pinned(engine, align_obj, vlist):
    """foreach_get("co") straight into page-locked memory, then icp_set_moving from it: the upload
    runs at the full host-link rate with no intermediate pageable copy."""
    n = len(align_obj.data.vertices)
    staging = PinnedArray((max(n, 1), 3), np.float32)
    try:
        coords = extract_coords(align_obj, out=staging.array.reshape(-1)[:3 * n])
        engine.set_moving(coords, vlist)          # returns after the copy has completed
    finally:
        staging.close()


class OBJECT_OT_align_pick_points(Operator):
    """Align two objects with 3 or more pairs of picked points (GPU solver)"""
    # the arithmetic of the reference's picked-points operator
    # (operators/align_pick_points.py:415-465).  Its split-view picking UI (invoke / modal / draw
    # callbacks, :106-357) is interactive bpy code and stays the addon's own; it fills
    # `align_points`, `base_points` (both in align-local space, :162,184) and `obj_align`, then calls
    # `align_obj`, which is what this class provides on top of icp_solve_pairs.
    bl_idname = "object.align_picked_points"
    bl_label = "Align: Picked Points"
    bl_options = {'REGISTER', 'UNDO'}

    def __init__(self, obj_align=None, align_points=(), base_points=()):
        self.obj_align = obj_align
        self.align_points = list(align_points)
        self.base_points = list(base_points)

    @classmethod
    def poll(cls, context):
        condition_1 = len(context.selected_objects) == 2
        condition_2 = context.object.type == 'MESH'
        return condition_1 and condition_2

    def align_obj(self, context):
        # unequal lists are cut to the shorter one (align_pick_points.py:417-422)
        if len(self.align_points) != len(self.base_points):
            if len(self.align_points) < len(self.base_points):
                self.base_points = self.base_points[0:len(self.align_points)]
            else:
                self.align_points = self.align_points[0:len(self.base_points)]
        k = len(self.base_points)
        A = np.zeros(shape=[3, k])
        B = np.zeros(shape=[3, k])
        for i in range(k):
            V1, V2 = self.align_points[i], self.base_points[i]
            A[0][i], A[1][i], A[2][i] = V1[0], V1[1], V1[2]
            B[0][i], B[1][i], B[2][i] = V2[0], V2[1], V2[2]
        settings = get_addon_preferences()
        # ValueError("input arrays are of wrong shape or type") below 3 pairs, like the reference's solver
        M = affine_matrix_from_points(A, B, shear=False, scale=(settings.align_meth == '1'), usesvd=True)
        new_mat = M.astype(np.float32)                                        # :446-449, fp64 -> fp32
        self.obj_align.matrix_world = self.obj_align.matrix_world @ array_to_matrix_like(
            new_mat, self.obj_align.matrix_world)                             # :453
        if settings.take_m_with:
            _apply_to_m_objects(context, [new_mat])                          # :455-459
        self.obj_align.update_tag()
        context.view_layer.update()
        self.M = M
        return M


class OBJECT_OT_icp_align(Operator):
    """Uses ICP alignment to iteratively align two objects (GPU engine)"""
    bl_idname = "object.align_icp"
    bl_label = "ICP Align"
    bl_options = {'REGISTER', 'UNDO'}

    # engine knobs that have no counterpart in the reference's preferences
    device = 0
    reject_fraction = 0.0
    correspondence = None      # None -> functions.DEFAULT_CORRESPONDENCE ("auto": surface if the base has faces)

    @classmethod
    def poll(cls, context):
        condition_1 = len(context.selected_objects) == 2
        condition_2 = context.object.type == 'MESH'
        return condition_1 and condition_2

    def execute(self, context):
        settings = get_addon_preferences()
        start = time.time()
        align_obj = context.object
        base_obj = [obj for obj in context.selected_objects if obj != align_obj][0]
        align_obj.rotation_mode = 'QUATERNION'

        vlist = build_vlist(align_obj)
        factor = stride_from_sample_fraction(settings.sample_fraction)   # ZeroDivisionError at 0
        thresh = settings.min_start
        if not thresh > 0:
            # the reference's make_pairs returns None here and the unpack at icp_align.py:101 raises
            raise TypeError("cannot unpack non-iterable NoneType object")

        base_nn = base_index_from_object(
            base_obj, context.evaluated_depsgraph_get(), self.correspondence, device=self.device, thresh=thresh,
            target_d=settings.target_d, use_target=settings.use_target,
            align_meth=settings.align_meth, stride=factor, reject_fraction=self.reject_fraction)
        try:
            eng = base_nn.engine
            upload_moving_pinned(eng, align_obj, vlist)
            eng.set_matrices(matrix_to_array(align_obj.matrix_world),
                             matrix_to_array(base_obj.matrix_world))
            try:
                done, converged, recs = eng.run(settings.icp_iterations)
            except IcpError as e:
                if e.code == N.ICP_ERR_TOO_FEW_PAIRS:
                    raise ValueError("input arrays are of wrong shape or type") from None
                raise
            align_obj.matrix_world = array_to_matrix_like(eng.matrix_world(), align_obj.matrix_world)
        finally:
            base_nn.close()
        if settings.take_m_with:
            _apply_to_m_objects(context, [r.M.astype(np.float32) for r in recs])
        align_obj.update_tag()
        context.view_layer.update()

        self.iterations = done
        self.converged = converged
        self.records = recs
        if settings.use_target and recs:
            last = recs[-1]
            print('Converged in %s iterations' % done if converged else 'Maxed out iterations')
            print('Final Translation: %f ' % last.trans)
            print('Final Avg Dist: %f' % last.mean_dist)
            print('Final St Dev %f' % last.std_dist)
            print('Avg last 5 rotation angle: %f' % np.mean([r.angle for r in recs[-5:]]))
        print('Aligned obj in %f sec' % (time.time() - start))
        return {'FINISHED'}


class OBJECT_OT_icp_align_feedback(Operator):
    """Uses ICP alignment to iteratively align two objects and redraws every n iterations (GPU engine)"""
    # Row N3 of SURVEY.md section 8f: the modal variant (reference operators/icp_align_feedback.py:31-304).
    # Same engine handle for the whole run; every TIMER tick runs `redraw_frequency` iterations
    # device-resident (icp_run) and writes matrix_world back once, so the viewport shows progress.
    # Two reference bugs are fixed knowingly: `take_m_with` is an unqualified name in iterate()
    # (icp_align_feedback.py:267, NameError) and the loop test `total_iters <= iters` (:113) runs one
    # iteration too many; here exactly `icp_iterations` iterations run, like the non-modal operator.
    bl_idname = "object.align_icp_redraw"
    bl_label = "ICP Align Redraw"
    bl_options = {'REGISTER', 'UNDO'}

    device = 0
    reject_fraction = 0.0
    correspondence = None
    _timer = None

    @classmethod
    def poll(cls, context):
        condition_1 = len(context.selected_objects) == 2
        condition_2 = context.object and context.object.type == 'MESH'
        return bool(condition_1 and condition_2)

    def invoke(self, context, event):
        wm = context.window_manager
        self._timer = wm.event_timer_add(time_step=0.01, window=context.window)   # :49
        wm.modal_handler_add(self)
        settings = get_addon_preferences()
        self.start = time.time()
        self.align_obj = context.object
        self.base_obj = [obj for obj in context.selected_objects if obj != self.align_obj][0]
        self.align_obj.rotation_mode = 'QUATERNION'
        self.vlist = build_vlist(self.align_obj)
        self.iters = settings.icp_iterations
        self.redraw_frequency = settings.redraw_frequency
        self.take_m_with = settings.take_m_with
        self.use_target = settings.use_target
        thresh = settings.min_start
        if not thresh > 0:
            raise TypeError("cannot unpack non-iterable NoneType object")
        self.base_nn = base_index_from_object(
            self.base_obj, context.evaluated_depsgraph_get(), self.correspondence, device=self.device, thresh=thresh,
            target_d=settings.target_d, use_target=settings.use_target, align_meth=settings.align_meth,
            stride=stride_from_sample_fraction(settings.sample_fraction),
            reject_fraction=self.reject_fraction)
        eng = self.base_nn.engine
        upload_moving_pinned(eng, self.align_obj, self.vlist)
        eng.set_matrices(matrix_to_array(self.align_obj.matrix_world),
                         matrix_to_array(self.base_obj.matrix_world))
        self.total_iters = 0
        self.converged = False
        self.records = []
        return {'RUNNING_MODAL'}

    def modal(self, context, event):
        if event.type in {'RIGHTMOUSE', 'ESC'}:
            return self.cancel(context)
        if event.type == 'TIMER':
            context.area.tag_redraw()
            if self.total_iters < self.iters and not self.converged:
                self.iterate(context, min(self.redraw_frequency, self.iters - self.total_iters))
                if self.total_iters < self.iters and not self.converged:
                    return {'RUNNING_MODAL'}
            return self.finish(context)
        return {'PASS_THROUGH'}

    def iterate(self, context, count=1):
        """`count` loop bodies of icp_align_feedback.py:250-288 in one device-resident batch."""
        eng = self.base_nn.engine
        try:
            done, converged, recs = eng.run(count)
        except IcpError as e:
            self._release(context)
            if e.code == N.ICP_ERR_TOO_FEW_PAIRS:
                raise ValueError("input arrays are of wrong shape or type") from None
            raise
        self.align_obj.matrix_world = array_to_matrix_like(eng.matrix_world(), self.align_obj.matrix_world)
        if self.take_m_with:
            _apply_to_m_objects(context, [r.M.astype(np.float32) for r in recs])
        self.align_obj.update_tag()
        self.total_iters += done
        self.converged = converged
        self.records.extend(recs)
        if converged and recs:
            print('Converged in %s iterations' % str(self.total_iters))
            print('Final Translation: %f ' % recs[-1].trans)
            print('Final Avg Dist: %f' % recs[-1].mean_dist)
            print('Final St Dev %f' % recs[-1].std_dist)

    def _release(self, context):
        if self._timer is not None:
            context.window_manager.event_timer_remove(self._timer)
            self._timer = None
        if getattr(self, "base_nn", None) is not None:
            self.base_nn.close()
            self.base_nn = None

    def finish(self, context):
        self._release(context)
        if self.use_target and not self.converged and self.records:
            print('Maxed out iterations')
            print('Final Avg Dist: %f' % self.records[-1].mean_dist)
            print('Final St Dev %f' % self.records[-1].std_dist)
        print('Aligned obj in %f sec' % (time.time() - self.start))
        return {'FINISHED'}

    def cancel(self, context):
        self._release(context)
        return {'CANCELLED'}
