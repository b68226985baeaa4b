"""Front end of oracle/_ref/glref — the reference's GLSL programs run by Mesa llvmpipe (a real OpenGL implementation) on an off-screen
context. TEST INFRASTRUCTURE ONLY (tests/test_gl_ref.py); see oracle/glref/glref_harness.cpp for what is the reference's own code there.

``render(scene)`` serialises a stillleben_b200.desc.SceneSpec into the harness' dump format, runs it in a subprocess (Mesa + LLVM stay
out of this interpreter) and returns the eight targets + the HDR buffer as numpy arrays shaped like tests/oracle_util.render's.
"""
import ctypes as C
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

from stillleben_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

GLREF = os.path.join(ROOT, "oracle", "_ref", "glref")
SHADER_DIR = os.path.join(build_ref.REF, "src", "shaders")


def available():
    """None when the harness can run here, else the reason it cannot."""
    if not os.path.exists(GLREF):
        return "oracle/_ref/glref not built (python oracle/build_ref.py gl)"
    if build_ref.find_mesa_libgl() is None:
        return "no Mesa libGL in this image (" + build_ref.MESA_LIBGL_GLOB + ")"
    return None


def _cm(m):
    return np.ascontiguousarray(np.asarray(m, np.float32).reshape(4, 4).T).tobytes()   # row-major m[r, c] -> column-major


def scene_lights(scene):
    """(directions, colours) the render pass uses: the light map's first three when there is one (render_pass.cpp:426-439), else the scene's."""
    if scene.light_map is None:
        return np.asarray(scene.light_directions, np.float32).reshape(3, 3), np.asarray(scene.light_colors, np.float32).reshape(3, 3)
    ld, lc = np.zeros((3, 3), np.float32), np.zeros((3, 3), np.float32)
    n = min(3, len(scene.light_map.light_directions))
    ld[:n] = np.asarray(scene.light_map.light_directions, np.float32).reshape(-1, 3)[:n]
    lc[:n] = np.asarray(scene.light_map.light_colors, np.float32).reshape(-1, 3)[:n]
    return ld, lc


def oracle_shadow_matrices(scene):
    """The oracle's frustum corners + shadow matrix per active light (render_pass.cpp:69-211; pinned on the reference's own two
    functions by tests/test_oracle_ref.py) — handed to the GL harness so that both sides use the SAME matrices."""
    import oracle_util as ou
    L = C.CDLL(ou.ORACLE_SO)
    fp = C.POINTER(C.c_float)
    L.orc_test_shadow_setup.argtypes = [fp, fp, C.c_int, fp, fp, fp, fp, fp, fp, fp]
    objs = list(scene.objects)        # computeFrustumCorners / computeShadowMapMatrix walk scene.objects(): the draw predicate does not apply
    n = len(objs)
    as_f = lambda rows, k: np.ascontiguousarray(np.concatenate(rows).astype(np.float32) if rows else np.zeros(k, np.float32))
    poses = as_f([np.asarray(o.pose, np.float32).T.reshape(-1) for o in objs], 16)
    pres = as_f([np.asarray(o.pretransform, np.float32).T.reshape(-1) for o in objs], 16)
    bmin = as_f([np.asarray(o.mesh.bbox_min, np.float32) for o in objs], 3)
    bmax = as_f([np.asarray(o.mesh.bbox_max, np.float32) for o in objs], 3)
    Pc = np.ascontiguousarray(np.asarray(scene.projection, np.float32).T.reshape(-1))
    Vc = np.ascontiguousarray(np.asarray(scene.world_to_cam, np.float32).T.reshape(-1))
    ld, lc = scene_lights(scene)
    active, mats = [], []
    for i in range(3):
        on = bool(np.any(lc[i] != 0) and np.any(ld[i] != 0))          # render_pass.cpp:433
        sm = np.zeros(16, np.float32)
        if on:
            corners, d = np.zeros(24, np.float32), np.ascontiguousarray(ld[i])
            L.orc_test_shadow_setup(*(a.ctypes.data_as(fp) for a in (Pc, Vc)), n, *(a.ctypes.data_as(fp) for a in (poses, pres, bmin, bmax, d, corners, sm)))
        active.append(int(on))
        mats.append(sm)                                                 # already column-major
    return active, mats


REFERENCE_LIGHTMAP_SIZES = (512, 32, 128, 512)     # environment cube, irradiance, prefilter, BRDF LUT (src/light_map.cpp:378,451,515,572)


def dump(scene, path, peel=None, lightmap_sizes=REFERENCE_LIGHTMAP_SIZES):
    textures, tex_index = [], {}

    def tex_of(img):
        if img is None:
            return -1
        if id(img) not in tex_index:
            tex_index[id(img)] = len(textures)
            textures.append(img)
        return tex_index[id(img)]

    meshes, mesh_index = [], {}
    for o in scene.objects:
        if id(o.mesh) not in mesh_index:
            mesh_index[id(o.mesh)] = len(meshes)
            meshes.append(o.mesh)
    active, shadow = oracle_shadow_matrices(scene)
    ld, lc = scene_lights(scene)
    out = [struct.pack("<IIii", 0x46524C47, 2, scene.width, scene.height), _cm(scene.projection), _cm(scene.world_to_cam),
           ld.reshape(9).tobytes(), lc.reshape(9).tobytes(), np.asarray(scene.ambient_light, np.float32).reshape(3).tobytes(), struct.pack("<3i", *active)]
    out += [m.tobytes() for m in shadow]
    out += [np.asarray(scene.background_plane_size, np.float32).reshape(2).tobytes(), _cm(scene.background_plane_pose),
            struct.pack("<ii", tex_of(scene.background_plane_texture), tex_of(scene.background_image)),
            struct.pack("<fi", float(scene.manual_exposure), int(bool(scene.ssao_enabled)))]
    lm = scene.light_map
    if lm is None:
        out.append(struct.pack("<i", 0))
    else:
        eq = np.ascontiguousarray(lm.equirect, np.float32)
        n = min(3, len(lm.light_directions))
        dirs, cols = np.zeros(9, np.float32), np.zeros(9, np.float32)
        dirs[:3 * n] = np.asarray(lm.light_directions, np.float32).reshape(-1)[:3 * n]
        cols[:3 * n] = np.asarray(lm.light_colors, np.float32).reshape(-1)[:3 * n]
        out += [struct.pack("<iii", 1, eq.shape[1], eq.shape[0]), eq.tobytes(), struct.pack("<i", n), dirs.tobytes(), cols.tobytes(),
                struct.pack("<4i", *lightmap_sizes)]
    if peel is None:
        out.append(struct.pack("<i", 0))
    else:
        out += [struct.pack("<i", 1), np.ascontiguousarray(peel, np.float32).tobytes()]
    mesh_blobs = []
    for m in meshes:
        blob = [struct.pack("<iiii", len(m.vertices), len(m.indices), len(m.submeshes), len(m.materials)), np.ascontiguousarray(m.vertices).tobytes(),
                np.ascontiguousarray(m.indices, np.uint32).tobytes()]
        blob += [struct.pack("<iii", int(o_), int(c_), int(mat)) for o_, c_, mat in m.submeshes]
        for mat in m.materials:
            slots = [mat.tex_base_color, mat.tex_normal, mat.tex_metallic_roughness, mat.tex_emissive, mat.tex_occlusion]
            blob.append(struct.pack("<4f4fff5i", *[float(v) for v in mat.base_color], *[float(v) for v in mat.emissive], float(mat.metallic), float(mat.roughness),
                                    *[tex_of(m.images[t]) if t >= 0 else -1 for t in slots]))
        mesh_blobs.append(b"".join(blob))
    obj_blobs = []
    for k, o in enumerate(scene.objects):
        obj_blobs.append(b"".join([
            struct.pack("<i", mesh_index[id(o.mesh)]), _cm(o.pose), _cm(o.pretransform),
            struct.pack("<iiffiii", int(o.class_index), int(o.instance_index) if o.instance_index else k + 1, float(o.metallic), float(o.roughness),
                        int(bool(o.casts_shadows)), int(bool(o.visible)), tex_of(o.sticker_texture)),
            _cm(o.sticker_projection), np.asarray(o.sticker_range, np.float32).reshape(4).tobytes()]))
    out.append(struct.pack("<i", len(textures)))
    for img in textures:
        px = np.ascontiguousarray(img.pixels, np.uint8)
        out += [struct.pack("<8i", px.shape[1], px.shape[0], px.shape[2], img.wrap_s, img.wrap_t, img.min_filter, img.mag_filter, img.kind), px.tobytes()]
    out.append(struct.pack("<i", len(meshes)))
    out += mesh_blobs
    out.append(struct.pack("<i", len(obj_blobs)))
    out += obj_blobs
    with open(path, "wb") as f:
        f.write(b"".join(out))


def gl_env(extra=None):
    """Environment of a glref process: the Mesa libGL, the fake Xlib next to the binary, the advertised GL / GLSL version; the shader
    files are read from the reference tree when it is here (a harness built earlier stays current) and from the copy compiled into
    the binary otherwise."""
    e = dict(os.environ, GLREF_LIBGL=build_ref.find_mesa_libgl(), MESA_GL_VERSION_OVERRIDE="4.5", MESA_GLSL_VERSION_OVERRIDE="450",
             LD_LIBRARY_PATH=os.path.join(os.path.dirname(GLREF), "glx") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    if os.path.isdir(SHADER_DIR):
        e["GLREF_SHADER_DIR"] = SHADER_DIR
    e.update(extra or {})
    return e


def render(scene, peel=None, env=None, lightmap_sizes=REFERENCE_LIGHTMAP_SIZES):
    """-> dict named like oracle_util.render's: the eight targets (abi.TARGET_NAMES) + 'hdr' (+ 'lightmap': env level 0 [6,e,e,4],
    irradiance [6,i,i,4], prefilter (packed levels), LUT [l,l,4] — the layout of Context.read_lightmap — when the scene has one)."""
    why = available()
    if why:
        raise RuntimeError(why)
    H, W = scene.height, scene.width
    with tempfile.TemporaryDirectory() as tmp:
        src, dst = os.path.join(tmp, "scene.bin"), os.path.join(tmp, "out.bin")
        dump(scene, src, peel, lightmap_sizes)
        e = gl_env(env)
        p = subprocess.run([GLREF, src, dst], env=e, capture_output=True, text=True, timeout=600)
        if p.returncode != 0:
            raise RuntimeError("glref failed (%d): %s" % (p.returncode, p.stderr[-4000:]))
        raw = np.fromfile(dst, np.uint8)
    outs, at = {}, 0
    for t, (dt, ch) in enumerate(abi.TARGET_FORMATS):
        n = H * W * ch * np.dtype(dt).itemsize
        outs[abi.TARGET_NAMES[t]] = raw[at:at + n].view(dt).reshape(H, W, ch).copy()
        at += n
    outs["hdr"] = raw[at:at + H * W * 16].view(np.float32).reshape(H, W, 4).copy()
    at += H * W * 16
    if scene.light_map is not None:
        e, i, pr, l = lightmap_sizes
        def take(n):
            nonlocal at
            a = raw[at:at + 4 * n].view(np.float32).copy()
            at += 4 * n
            return a
        env0 = take(6 * e * e * 4).reshape(6, e, e, 4)
        irr = take(6 * i * i * 4).reshape(6, i, i, 4)
        pre = take(sum(6 * (pr >> m) ** 2 * 4 for m in range(5)))
        lut = take(l * l * 4).reshape(l, l, 4)
        outs["lightmap"] = (env0, irr, pre, lut)
    if at < raw.size:                   # GLREF_DUMP_SHADOW: the shadow-map array, float depth
        outs["shadow"] = raw[at:at + 3 * 2048 * 2048 * 4].view(np.float32).reshape(3, 2048, 2048).copy()
    outs["stderr"] = p.stderr
    return outs
