"""ctypes mirror of include/slb.h (struct layouts + prototypes).

Only declarations live here — no rendering logic. The same struct layouts are used by the tests to
talk to the CPU oracle (whose handles are plain pointers too), which is why the handle fields are
``c_void_p`` instead of typed pointers.
"""
import ctypes as C

import numpy as np

SLB_ABI_VERSION = 1
NUM_LIGHTS = 3
VERTEX_STRIDE = 68
SHADOW_RES = 2048
INVALID_COORD = 3000.0

OK, ERR_INVALID_ARGUMENT, ERR_RUNTIME, ERR_CUDA, ERR_OUT_OF_MEMORY = range(5)

(TARGET_RGB, TARGET_COORD, TARGET_CLASS, TARGET_INSTANCE, TARGET_NORMAL, TARGET_VERTEX_INDEX, TARGET_BARY,
 TARGET_CAM_COORD) = range(8)
NUM_TARGETS = 8
TARGETS_SIX = 0x1F
TARGETS_ALL = 0xFF

TARGET_NAMES = ["rgb", "coord", "class_index", "instance_index", "normals", "vertex_index", "barycentric", "cam_coord"]
# (numpy dtype, channels) of every target, in attachment order (reference: src/render_pass.cpp:347-365)
TARGET_FORMATS = [
    (np.uint8, 4), (np.float32, 4), (np.uint16, 1), (np.uint16, 1), (np.float32, 4), (np.uint32, 4), (np.float32, 4),
    (np.float32, 4),
]
TARGET_BYTES_PER_PIXEL = [4, 16, 2, 2, 16, 16, 16, 16]

WRAP_REPEAT, WRAP_CLAMP_TO_EDGE, WRAP_MIRRORED_REPEAT, WRAP_CLAMP_TO_BORDER = range(4)
(FILTER_NEAREST, FILTER_LINEAR, FILTER_NEAREST_MIPMAP_NEAREST, FILTER_LINEAR_MIPMAP_NEAREST,
 FILTER_NEAREST_MIPMAP_LINEAR, FILTER_LINEAR_MIPMAP_LINEAR) = range(6)
TEXTURE_2D, TEXTURE_RECT = 0, 1

OPT_TIME_KERNELS, OPT_KEEP_HDR, OPT_MAX_SUBBATCH, OPT_DIRECT_MAX, OPT_WARP_MAX, OPT_LEAN_SHADE, OPT_HUGE_IN_SHADE = 1, 2, 3, 4, 5, 6, 7
OPT_HUGE_PREPARE, OPT_SHADOW_MASK, OPT_OVERLAP = 8, 9, 10

# numpy dtype of the 68-byte consolidated vertex (reference: src/mesh_tools/consolidate.cpp:53-61)
VERTEX_DTYPE = np.dtype([
    ("position", np.float32, 3),
    ("uv", np.float32, 2),
    ("color", np.float32, 4),
    ("tangent", np.float32, 4),
    ("vertex_index", np.uint32),
    ("normal", np.float32, 3),
])
assert VERTEX_DTYPE.itemsize == VERTEX_STRIDE

Mat4 = C.c_float * 16
Vec3 = C.c_float * 3


class Image(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("wrap_s", C.c_int32), ("wrap_t", C.c_int32), ("min_filter", C.c_int32), ("mag_filter", C.c_int32)]


class Submesh(C.Structure):
    _fields_ = [("index_offset", C.c_uint32), ("index_count", C.c_uint32), ("material", C.c_int32),
                ("reserved", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("base_color", C.c_float * 4), ("emissive", C.c_float * 4), ("metallic", C.c_float),
                ("roughness", C.c_float), ("tex_base_color", C.c_int32), ("tex_normal", C.c_int32),
                ("tex_metallic_roughness", C.c_int32), ("tex_emissive", C.c_int32), ("tex_occlusion", C.c_int32),
                ("reserved", C.c_int32)]


class LightmapDesc(C.Structure):
    _fields_ = [("equirect_rgb", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("n_lights", C.c_int32),
                ("light_directions", (C.c_float * 3) * NUM_LIGHTS), ("light_colors", (C.c_float * 3) * NUM_LIGHTS)]


class ObjectDesc(C.Structure):
    _fields_ = [("mesh", C.c_void_p), ("pose", Mat4), ("pretransform", Mat4), ("class_index", C.c_uint32),
                ("instance_index", C.c_uint32), ("metallic", C.c_float), ("roughness", C.c_float),
                ("casts_shadows", C.c_int32), ("visible", C.c_int32), ("sticker_texture", C.c_void_p),
                ("sticker_projection", Mat4), ("sticker_range", C.c_float * 4)]


class SceneDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("projection", Mat4), ("world_to_cam", Mat4),
                ("light_directions", (C.c_float * 3) * NUM_LIGHTS), ("light_colors", (C.c_float * 3) * NUM_LIGHTS),
                ("ambient_light", C.c_float * 3), ("light_map", C.c_void_p),
                ("background_plane_size", C.c_float * 2), ("background_plane_pose", Mat4),
                ("background_plane_texture", C.c_void_p), ("background_image", C.c_void_p),
                ("manual_exposure", C.c_float), ("ssao_enabled", C.c_int32), ("objects", C.POINTER(ObjectDesc)),
                ("n_objects", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("frames_rendered", C.c_uint64), ("triangles_submitted", C.c_uint64),
                ("triangles_binned", C.c_uint64), ("bytes_h2d", C.c_uint64), ("bytes_d2h", C.c_uint64),
                ("last_kernel_ms", C.c_float * 8)]


def mat4_to_c(m):
    """4x4 row-major (numpy/torch convention, m[r, c]) -> column-major float[16] (Magnum order)."""
    a = np.asarray(m, dtype=np.float32).reshape(4, 4)
    return Mat4(*a.T.reshape(-1).tolist())


# every symbol include/slb.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "slb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "slb_ctx_destroy": (None, [C.c_void_p]),
    "slb_last_error": (C.c_char_p, [C.c_void_p]),
    "slb_abi_version": (C.c_int, []),
    "slb_ctx_device": (C.c_int, [C.c_void_p]),
    "slb_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "slb_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "slb_host_free": (None, [C.c_void_p, C.c_void_p]),
    "slb_mesh_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(Submesh),
                                   C.c_uint32, C.POINTER(Material), C.c_uint32, C.POINTER(Image), C.c_uint32,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_void_p)]),
    "slb_mesh_update_vertices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "slb_mesh_update_positions_and_colors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "slb_mesh_set_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "slb_mesh_set_colors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "slb_mesh_recompute_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "slb_mesh_read_vertices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "slb_mesh_destroy": (None, [C.c_void_p, C.c_void_p]),
    "slb_texture_create": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int, C.POINTER(C.c_void_p)]),
    "slb_texture_destroy": (None, [C.c_void_p, C.c_void_p]),
    "slb_texture_read_level": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                          C.c_void_p]),
    "slb_lightmap_create": (C.c_int, [C.c_void_p, C.POINTER(LightmapDesc), C.POINTER(C.c_void_p)]),
    "slb_lightmap_create_from_maps": (C.c_int, [C.c_void_p, C.POINTER(LightmapDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "slb_lightmap_create_ex": (C.c_int, [C.c_void_p, C.POINTER(LightmapDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_void_p)]),
    "slb_lightmap_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "slb_lightmap_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "slb_lightmap_destroy": (None, [C.c_void_p, C.c_void_p]),
    "slb_result_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p)]),
    "slb_result_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "slb_result_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]),
    "slb_result_read_hdr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]),
    "slb_result_destroy": (None, [C.c_void_p, C.c_void_p]),
    "slb_render_batch": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_void_p]),
    "slb_render_batch_host": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc), C.c_int32, C.c_uint32,
                                         C.POINTER(C.c_void_p)]),
    "slb_ctx_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "slb_ctx_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "slb_diff_sobel_valid_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                             C.c_void_p]),
    "slb_diff_dilate_object_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                               C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "slb_camera_model": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p]),
    "slb_png_bound": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "slb_png_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p]),
    "slb_jpeg_bound": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "slb_jpeg_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p]),
    "slb_diff_pose_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
}


CAM_CHROMATIC, CAM_BLUR, CAM_EXPOSURE, CAM_NOISE, CAM_CLAMP, CAM_HUE, CAM_POST_BLUR, CAM_ALL = 1, 2, 4, 8, 16, 32, 64, 127


class CameraParams(C.Structure):
    _fields_ = [("chromatic_translation", (C.c_float * 2) * 3), ("chromatic_scaling", C.c_float * 3), ("blur_sigma", C.c_float),
                ("exposure_deltaS", C.c_float), ("do_noise", C.c_int32), ("noise_a", C.c_float), ("noise_b", C.c_float),
                ("hue_shift", C.c_float), ("stages", C.c_uint32), ("seed", C.c_uint64)]


def bind(lib):
    """Attach restype/argtypes for every slb_* symbol; raises AttributeError if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
