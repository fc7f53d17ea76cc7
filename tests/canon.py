"""Test helper: canonical-layout numpy copies of a scene + driver of the CPU emulation harness (tests/emul)."""
import ctypes as C
import os
import subprocess

import numpy as np

from deodr_b200 import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))


def canonical_arrays(scene):
    a = {
        "faces": np.ascontiguousarray(scene.faces, dtype=np.uint32),
        "faces_uv": np.ascontiguousarray(scene.faces_uv, dtype=np.uint32),
        "ij": np.ascontiguousarray(scene.ij, dtype=np.float64),
        "depths": np.ascontiguousarray(scene.depths, dtype=np.float64),
        "uv": np.ascontiguousarray(scene.uv, dtype=np.float64),
        "colors": np.ascontiguousarray(scene.colors, dtype=np.float32),
        "shade": np.ascontiguousarray(scene.shade, dtype=np.float32),
        "edgeflags": np.ascontiguousarray(scene.edgeflags, dtype=np.uint8),
        "textured": np.ascontiguousarray(scene.textured, dtype=np.uint8),
        "shaded": np.ascontiguousarray(scene.shaded, dtype=np.uint8),
        "texture": np.ascontiguousarray(scene.texture, dtype=np.float32),
    }
    if scene.background_image is not None:
        a["background_image"] = np.ascontiguousarray(scene.background_image, dtype=np.float32)
    else:
        a["background_color"] = np.ascontiguousarray(scene.background_color, dtype=np.float32)
    return a


def view_of(scene, arrays, ptr=lambda arr: arr.ctypes.data):
    v = _cabi.SceneView()
    for name in ("faces", "faces_uv", "ij", "depths", "uv", "colors", "shade", "edgeflags", "textured", "shaded",
                 "texture"):
        setattr(v, name, ptr(arrays[name]))
    v.background_image = ptr(arrays["background_image"]) if "background_image" in arrays else None
    v.background_color = ptr(arrays["background_color"]) if "background_color" in arrays else None
    v.nb_triangles = arrays["faces"].shape[0]
    v.nb_vertices = arrays["depths"].shape[0]
    v.nb_uv = arrays["uv"].shape[0]
    v.height, v.width, v.nb_colors = int(scene.height), int(scene.width), int(scene.nb_colors)
    v.texture_height, v.texture_width = arrays["texture"].shape[0], arrays["texture"].shape[1]
    v.clockwise = int(bool(scene.clockwise))
    v.backface_culling = int(bool(scene.backface_culling))
    v.strict_edge = int(bool(scene.strict_edge))
    v.perspective_correct = int(bool(scene.perspective_correct))
    v.integer_pixel_centers = int(bool(scene.integer_pixel_centers))
    return v


class Emulator:
    def __init__(self, exact_short_wrap: bool = False):
        """exact_short_wrap: the build with DEODR_EXACT_SHORT_WRAP=1 (rmath.h) instead of the default one."""
        name = "libemul_wrap.so" if exact_short_wrap else "libemul.so"
        subprocess.run(["make", "-C", os.path.join(_HERE, "emul"), name], check=True, capture_output=True)
        self.lib = C.CDLL(os.path.join(_HERE, "emul", name))
        self.lib.emul_render.argtypes = [C.POINTER(_cabi.SceneView), C.c_double] + [C.c_void_p] * 4
        self.lib.emul_weight_scale.argtypes = [C.POINTER(_cabi.SceneView), C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.emul_weight_scale.restype = None
        self.lib.emul_render_error.argtypes = [C.POINTER(_cabi.SceneView), C.c_double] + [C.c_void_p] * 6
        self.lib.emul_render_error_b.argtypes = [C.POINTER(_cabi.SceneView), C.c_double] + [C.c_void_p] * 5 + [
            C.POINTER(_cabi.Grads), C.c_int]
        self.lib.emul_render_b.argtypes = [C.POINTER(_cabi.SceneView), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(_cabi.Grads)]

    def render(self, scene, sigma):
        a = canonical_arrays(scene)
        v = view_of(scene, a)
        H, W, Cc = scene.height, scene.width, scene.nb_colors
        out = {
            "image": np.zeros((H, W, Cc), np.float32), "z": np.zeros((H, W)), "owner": np.zeros((H, W), np.int32),
            "face_id": np.zeros((H, W), np.int32),
        }
        rc = self.lib.emul_render(C.byref(v), float(sigma), out["image"].ctypes.data, out["z"].ctypes.data,
                                  out["owner"].ctypes.data, out["face_id"].ctypes.data)
        assert rc == 0
        out["_arrays"], out["_view"] = a, v
        out["weight_scale"] = np.ones((H, W), np.float32)
        self.lib.emul_weight_scale(C.byref(v), out["face_id"].ctypes.data, out["z"].ctypes.data,
                                   out["weight_scale"].ctypes.data)
        out["ties"] = self.lib.emul_num_ties()
        out["edges"] = self.lib.emul_num_edges()
        out["tri_refs"] = self.lib.emul_tri_refs()
        return out

    def set_record_rows(self, rows):
        """Height limit of the record path for triangles that are not small (the device: 16, or 0 for tiny scenes)."""
        self.lib.emul_set_record_rows(int(rows))

    def set_small_textured(self, on):
        """May textured triangles take the triangle-parallel adjoint (TriBins::small_textured)?"""
        self.lib.emul_set_small_textured(int(bool(on)))

    def set_small_records(self, on):
        """Adjoint of the small triangles: triangle-parallel (the device's default) or record-parallel (k_small_rec_bwd)."""
        self.lib.emul_set_small_records(int(bool(on)))

    def build_plan(self, scene, sigma):
        """Segment capacities from `scene` (count-only pass + scans), kept for the render_planned calls that follow."""
        a = canonical_arrays(scene)
        v = view_of(scene, a)
        self.lib.emul_build_plan.argtypes = [C.POINTER(_cabi.SceneView), C.c_double]
        assert self.lib.emul_build_plan(C.byref(v), float(sigma)) == 0

    def render_planned(self, scene, sigma):
        """Forward pass of `scene` against the plan of an EARLIER scene of the same shape -> (verdict word, outputs)."""
        a = canonical_arrays(scene)
        v = view_of(scene, a)
        H, W, Cc = scene.height, scene.width, scene.nb_colors
        out = {
            "image": np.zeros((H, W, Cc), np.float32), "z": np.zeros((H, W)), "owner": np.zeros((H, W), np.int32),
            "face_id": np.zeros((H, W), np.int32),
        }
        self.lib.emul_render_planned.argtypes = [C.POINTER(_cabi.SceneView), C.c_double] + [C.c_void_p] * 4
        verdict = self.lib.emul_render_planned(C.byref(v), float(sigma), out["image"].ctypes.data, out["z"].ctypes.data,
                                               out["owner"].ctypes.data, out["face_id"].ctypes.data)
        out["_arrays"], out["_view"] = a, v
        return verdict, out

    def render_b(self, scene, sigma, fwd, image_b):
        a = fwd["_arrays"]
        ib = np.ascontiguousarray(image_b, dtype=np.float32)
        g = {
            "ij_b": np.zeros(a["ij"].shape, np.float32), "colors_b": np.zeros(a["colors"].shape, np.float32),
            "uv_b": np.zeros(a["uv"].shape, np.float32), "shade_b": np.zeros(a["shade"].shape, np.float32),
            "texture_b": np.zeros(a["texture"].shape, np.float32),
        }
        gs = _cabi.Grads(*(g[k].ctypes.data for k in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b")))
        rc = self.lib.emul_render_b(C.byref(fwd["_view"]), float(sigma), fwd["z"].ctypes.data, fwd["owner"].ctypes.data,
                                    ib.ctypes.data, C.byref(gs))
        assert rc == 0
        return g

    # ---- antialiase_error mode (row f3) ----
    def render_error(self, scene, sigma, obs):
        a = canonical_arrays(scene)
        v = view_of(scene, a)
        H, W, Cc = scene.height, scene.width, scene.nb_colors
        out = {
            "image": np.zeros((H, W, Cc), np.float32), "z": np.zeros((H, W)), "owner": np.zeros((H, W), np.int32),
            "face_id": np.zeros((H, W), np.int32), "err": np.zeros((H, W), np.float32),
            "obs": np.ascontiguousarray(obs, dtype=np.float32),
        }
        rc = self.lib.emul_render_error(C.byref(v), float(sigma), out["obs"].ctypes.data, out["image"].ctypes.data,
                                        out["z"].ctypes.data, out["owner"].ctypes.data, out["face_id"].ctypes.data,
                                        out["err"].ctypes.data)
        assert rc == 0
        out["_arrays"], out["_view"] = a, v
        return out

    def render_error_b(self, scene, sigma, fwd, err_b, compat=True):
        a = fwd["_arrays"]
        eb = np.ascontiguousarray(err_b, dtype=np.float32)
        g = {
            "ij_b": np.zeros(a["ij"].shape, np.float32), "colors_b": np.zeros(a["colors"].shape, np.float32),
            "uv_b": np.zeros(a["uv"].shape, np.float32), "shade_b": np.zeros(a["shade"].shape, np.float32),
            "texture_b": np.zeros(a["texture"].shape, np.float32),
        }
        gs = _cabi.Grads(*(g[k].ctypes.data for k in ("ij_b", "colors_b", "uv_b", "shade_b", "texture_b")))
        rc = self.lib.emul_render_error_b(C.byref(fwd["_view"]), float(sigma), fwd["z"].ctypes.data,
                                          fwd["owner"].ctypes.data, fwd["image"].ctypes.data, fwd["obs"].ctypes.data,
                                          eb.ctypes.data, C.byref(gs), int(bool(compat)))
        assert rc == 0
        return g
