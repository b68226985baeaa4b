"""Recipe for oracle/_ref/: pieces of the REFERENCE ITSELF compiled from the sources where they lie under
/root/reference (never copied into this repo), used to pin the oracle restatement (tests/test_oracle_ref.py) and to
generate the committed golden fixtures (tests/golden/make_ref_golden.py). Test infrastructure only.

  diff      python/src/diff.cu + bridge_diff.cpp  -> oracle/_ref/diff/libstillleben_diff_python.so
            (a self-contained torch extension: torch/extension.h only; CPU and CUDA branches)
  meshtool  src/mesh_tools/{consolidate,compute_tangents}.cpp + GL-less Corrade/Magnum/CgltfImporter from contrib/
            -> oracle/_ref/meshtool (dumps the reference's own consolidated 68-byte vertex stream)
  glsl      src/shaders/*.{vert,frag,glsl} compiled VERBATIM as C++ behind oracle/glsl_compat.h
            -> oracle/_ref/libglslref.so
  host      the pure-math member functions behind the render pass's inputs — Scene::setCameraLookAt / setCameraIntrinsics /
            setCameraFromFOV / setCameraPose, Mesh::centerBBox / scaleToBBoxDiagonal / setPretransform / bbox,
            Object::stickerViewProjection — cut out of src/{scene,mesh,object}.cpp -> oracle/_ref/libhostref.so
  frame     computeFrustumCorners + computeShadowMapMatrix, extracted verbatim from src/render_pass.cpp at build time and
            compiled against the same GL-less Magnum -> oracle/_ref/libframeref.so

  gl        the reference's GLSL programs run by a real OpenGL implementation: oracle/glref/glref_harness.cpp (the enums, the #define
            header and every uniform setter of src/shaders/render_shader.cpp cut out and compiled in; the shader files read from
            /root/reference/src/shaders at run time) + a do-nothing Xlib (oracle/glref/fake_x11.c) under the Mesa llvmpipe libGL that
            ships with Nsight Compute in this image -> oracle/_ref/glref, oracle/_ref/glx/libX11.so.6, libXext.so.6

Usage: python oracle/build_ref.py [diff] [meshtool] [glsl] [frame] [host] [gl]   (no argument = everything that is not built yet)
Outputs only into oracle/_ref/ (git-ignored, not gpurun-ignored). Needs /root/reference; a no-op without it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SLB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build_diff(force=False):
    """The reference's torch extension for config 4 (python/src/diff.cu:13-193, bridge_diff.cpp:13-176)."""
    out = os.path.join(OUT, "diff")
    so = os.path.join(out, "libstillleben_diff_python.so")
    if os.path.exists(so) and not force:
        return so
    os.makedirs(out, exist_ok=True)
    from torch.utils import cpp_extension
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    cpp_extension.load(
        name="libstillleben_diff_python",
        sources=[os.path.join(REF, "python/src/bridge_diff.cpp"), os.path.join(REF, "python/src/diff.cu")],
        extra_include_paths=[os.path.join(REF, "python/src")],
        extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"],
        build_directory=out, is_python_module=False, verbose=False, with_cuda=True)
    return so


def load_diff():
    """Import oracle/_ref/diff/libstillleben_diff_python.so (the prebuilt file; /root/reference is not needed)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    so = os.path.join(OUT, "diff", "libstillleben_diff_python.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("libstillleben_diff_python", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_meshtool(force=False):
    """The reference's mesh front end: src/mesh_tools/{consolidate,compute_tangents}.cpp + vendored GL-less
    Corrade / Magnum / CgltfImporter (oracle/build_magnum.sh), driven by oracle/ref_meshtool.cpp."""
    exe = os.path.join(OUT, "meshtool")
    if os.path.exists(exe) and not force:
        return exe
    prefix = subprocess.check_output(["bash", os.path.join(HERE, "build_magnum.sh")], text=True).strip().splitlines()[-1]
    libs = [f"{prefix}/lib/magnum/importers/lib{n}.a" for n in ("CgltfImporter", "StbImageImporter", "AnyImageImporter")]
    libs += [f"{prefix}/lib/lib{n}.a" for n in ("MagnumMeshTools", "MagnumTrade", "Magnum", "CorradePluginManager", "CorradeUtility")]
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", f"-I{prefix}/include", f"-I{REF}/include", os.path.join(HERE, "ref_meshtool.cpp"),
                           os.path.join(REF, "src/mesh_tools/consolidate.cpp"), os.path.join(REF, "src/mesh_tools/compute_tangents.cpp"),
                           "-o", exe] + libs + ["-ldl"])
    return exe


GLSL_SOURCES = ["render_shader.glsl", "render_shader.vert", "render_shader.frag", "shadow_shader.vert", "tone_map_shader.frag",
                "ssao_shader.frag", "ssao_apply_shader.frag", "background_shader.vert", "background_shader.frag",
                "background_cube_shader.vert", "background_cube_shader.frag", "cubemap_shader_equirectangular.frag",
                "cubemap_shader_irradiance.frag", "cubemap_shader_prefilter.frag", "brdf_shader.frag"]


def build_glsl(force=False):
    """The reference's GLSL programs (src/shaders/*) compiled VERBATIM as C++: oracle/glsl_to_cpp.py -> oracle/_ref/gen/*.inc,
    oracle/glslref_harness.cpp + oracle/glsl_compat.h + the oracle sources -> oracle/_ref/libglslref.so."""
    so = os.path.join(OUT, "libglslref.so")
    if os.path.exists(so) and not force:
        return so
    gen = os.path.join(OUT, "gen")
    os.makedirs(gen, exist_ok=True)
    sys.path.insert(0, HERE)
    import glsl_to_cpp
    for name in GLSL_SOURCES:
        src = os.path.join(REF, "src/shaders", name)
        with open(src) as f:
            text = glsl_to_cpp.translate(f.read())
        with open(os.path.join(gen, name + ".inc"), "w") as f:
            f.write(f"// generated by oracle/glsl_to_cpp.py from {src} - do not commit\n" + text)
    srcs = [os.path.join(HERE, n) for n in ("glslref_harness.cpp", "orc_render.cpp", "orc_texture.cpp", "orc_lightmap.cpp", "orc_diff.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared",
                           f"-I{HERE}", "-o", so] + srcs)
    return so


def build_frame(force=False):
    """The two host-side functions of RenderPass::render that are pure Magnum math — computeFrustumCorners and
    computeShadowMapMatrix (src/render_pass.cpp, between `using FrustumCorners` and the end of the anonymous namespace) — cut out
    of the reference source at build time (the rest of the file needs GL) and compiled with oracle/ref_frame_harness.cpp."""
    so = os.path.join(OUT, "libframeref.so")
    if os.path.exists(so) and not force:
        return so
    gen = os.path.join(OUT, "gen")
    os.makedirs(gen, exist_ok=True)
    src = os.path.join(REF, "src/render_pass.cpp")
    lines = open(src).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith("using FrustumCorners"))
    end = next(i for i, ln in enumerate(lines) if ln.startswith("RenderPass::Result::Result"))
    while not lines[end - 1].startswith("}"):          # back over blank lines to the `}` that closes the anonymous namespace
        end -= 1
    text = "\n".join(lines[start:end - 1])
    assert "computeFrustumCorners" in text and "computeShadowMapMatrix" in text and "GL::" not in text
    with open(os.path.join(gen, "render_pass_frame.inc"), "w") as f:
        f.write(f"// cut from {src}:{start + 1}-{end - 1} by oracle/build_ref.py - do not commit\n" + text + "\n")
    prefix = subprocess.check_output(["bash", os.path.join(HERE, "build_magnum.sh")], text=True).strip().splitlines()[-1]
    libs = [f"{prefix}/lib/lib{n}.a" for n in ("Magnum", "CorradeUtility")]
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", f"-I{prefix}/include", f"-I{HERE}",
                           os.path.join(HERE, "ref_frame_harness.cpp"), "-o", so] + libs)
    return so


def _cut(path, first_prefix, stop_prefix):
    """Lines of `path` from the first line starting with `first_prefix` up to (not including) the first later line starting
    with `stop_prefix`, trailing blank lines dropped."""
    lines = open(path).read().splitlines()
    a = next(i for i, ln in enumerate(lines) if ln.startswith(first_prefix))
    b = next(i for i in range(a + 1, len(lines)) if lines[i].startswith(stop_prefix))
    while not lines[b - 1].strip():
        b -= 1
    return f"// cut from {path}:{a + 1}-{b} by oracle/build_ref.py - do not commit\n" + "\n".join(lines[a:b]) + "\n"


def build_host(force=False):
    """Host-side pure-math member functions of Scene / Mesh / Object (see the module docstring), compiled with
    oracle/ref_host_harness.cpp against the reference's GL-less Magnum."""
    so = os.path.join(OUT, "libhostref.so")
    if os.path.exists(so) and not force:
        return so
    gen = os.path.join(OUT, "gen")
    os.makedirs(gen, exist_ok=True)
    cuts = {"scene_camera.inc": (os.path.join(REF, "src/scene.cpp"), "void Scene::setCameraPose", "Magnum::Matrix4 Scene::projectionMatrix"),
            "mesh_pretransform.inc": (os.path.join(REF, "src/mesh.cpp"), "void Mesh::centerBBox", "void Mesh::setClassIndex"),
            "mesh_normals.inc": (os.path.join(REF, "src/mesh.cpp"), "void Mesh::recomputeNormals", "void Mesh::recompileMesh"),
            "mesh_vertex_edit.inc": (os.path.join(REF, "src/mesh.cpp"), "void Mesh::updateVertexPositionsAndColors", "void Mesh::setVertexColors"),
            "lightmap_specs.inc": (os.path.join(REF, "src/light_map.cpp"), "    struct IBLSpec", "    Containers::Optional<Magnum::GL::Texture2D> loadTexture"),
            "lightmap_lights.inc": (os.path.join(REF, "src/light_map.cpp"), "        auto addLight = [&]", "    }"),
            "ssao_tables.inc": (os.path.join(REF, "src/shaders/ssao_shader.cpp"), "    // Create noise texture", "SSAOShader& SSAOShader::bindCoordinates"),
            "object_sticker.inc": (os.path.join(REF, "src/object.cpp"), "Magnum::Matrix4 Object::stickerViewProjection", "void Object::setStatic")}
    for name, (path, first, stop) in cuts.items():
        text = _cut(path, first, stop)
        assert ("GL::" not in text or name == "ssao_tables.inc") and "physx" not in text.lower()
        with open(os.path.join(gen, name), "w") as f:
            f.write(text)
    prefix = subprocess.check_output(["bash", os.path.join(HERE, "build_magnum.sh")], text=True).strip().splitlines()[-1]
    libs = [f"{prefix}/lib/lib{n}.a" for n in ("Magnum", "CorradeUtility")]
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", f"-I{prefix}/include", f"-I{HERE}",
                           os.path.join(HERE, "ref_host_harness.cpp"), "-o", so] + libs)
    return so


MESA_LIBGL_GLOB = "/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_2_11_3-x64/Mesa/libGL.so.1"


def find_mesa_libgl():
    """The software OpenGL of this image: the Mesa 18.1 llvmpipe libGL bundled with Nsight Compute (needs libX11 / libXext, which the
    image does not have: oracle/glref/fake_x11.c stands in). None when absent."""
    import glob
    hits = sorted(glob.glob(os.environ.get("GLREF_LIBGL", MESA_LIBGL_GLOB)))
    return hits[-1] if hits else None


def build_gl(force=False):
    """oracle/_ref/glref (see the module docstring). The cuts out of src/shaders/render_shader.cpp: the TextureInput / Uniform enums
    (:34-75), the statements building the #define header (:95-214), every setter (:233-end)."""
    exe = os.path.join(OUT, "glref")
    if os.path.exists(exe) and not force:
        return exe
    gen, glx = os.path.join(OUT, "gen"), os.path.join(OUT, "glx")
    os.makedirs(gen, exist_ok=True)
    os.makedirs(glx, exist_ok=True)
    src = os.path.join(REF, "src/shaders/render_shader.cpp")
    lines = open(src).read().splitlines()
    first = next(i for i, ln in enumerate(lines) if ln.startswith("RenderShader& RenderShader::setTransformations"))
    cuts = {"render_shader_enums.inc": "namespace\n{\n" + _cut(src, "    enum class TextureInput", "}") + "}\n",
            "render_shader_header.inc": _cut(src, "    std::string header = ", "    vert.addSource(header)"),
            "render_shader_setters.inc": f"// cut from {src}:{first + 1}-{len(lines)} by oracle/build_ref.py - do not commit\n" + "\n".join(lines[first:]) + "\n"}
    cuts["ssao_tables.inc"] = _cut(os.path.join(REF, "src/shaders/ssao_shader.cpp"), "    // Create noise texture", "SSAOShader& SSAOShader::bindCoordinates")
    cuts["lightmap_cube.inc"] = _cut(os.path.join(REF, "src/light_map.cpp"), "    struct CubeMapSide", "}")
    shader_dir = os.path.join(REF, "src/shaders")
    blob = []
    for fn in sorted(os.listdir(shader_dir)):
        if fn.endswith((".vert", ".frag", ".geom", ".glsl")):
            text = open(os.path.join(shader_dir, fn)).read()
            assert ')GLSLFILE"' not in text
            blob.append('{"%s", R"GLSLFILE(%s)GLSLFILE"},\n' % (fn, text))
    cuts["shader_blob.inc"] = f"// {shader_dir}/* as string literals, by oracle/build_ref.py - do not commit\n" + "".join(blob)
    for name, text in cuts.items():
        assert "physx" not in text.lower()
        with open(os.path.join(gen, name), "w") as f:
            f.write(text)
    gl = os.path.join(HERE, "glref")
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-Wl,-soname,libX11.so.6", f"-I{gl}", os.path.join(gl, "fake_x11.c"), "-o", os.path.join(glx, "libX11.so.6")])
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-Wl,-soname,libXext.so.6", "-x", "c", "/dev/null", "-o", os.path.join(glx, "libXext.so.6")])
    prefix = subprocess.check_output(["bash", os.path.join(HERE, "build_magnum.sh")], text=True).strip().splitlines()[-1]
    libs = [f"{prefix}/lib/lib{n}.a" for n in ("MagnumPrimitives", "MagnumTrade", "Magnum", "CorradePluginManager", "CorradeUtility")]
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{prefix}/include", f"-I{HERE}", f"-I{gl}", os.path.join(gl, "glref_harness.cpp"), "-o", exe,
                           f"-L{glx}", "-l:libX11.so.6", "-Wl,-rpath,$ORIGIN/glx"] + libs + ["-ldl"])
    return exe


def main(argv):
    if not os.path.isdir(REF):
        print("oracle/build_ref.py: no reference tree at", REF, "- keeping prebuilt oracle/_ref as is")
        return 0
    want = [a for a in argv if not a.startswith("-")] or ["diff", "meshtool", "glsl", "frame", "host", "gl"]
    if "diff" in want:
        print("diff ->", build_diff(force="--force" in argv))
    if "meshtool" in want:
        print("meshtool ->", build_meshtool(force="--force" in argv))
    if "glsl" in want:
        print("glsl ->", build_glsl(force="--force" in argv))
    if "frame" in want:
        print("frame ->", build_frame(force="--force" in argv))
    if "host" in want:
        print("host ->", build_host(force="--force" in argv))
    if "gl" in want:
        print("gl ->", build_gl(force="--force" in argv))
    return 0


if __name__ == "__main__":
    sys.exit(main([a for a in sys.argv[1:]]))
