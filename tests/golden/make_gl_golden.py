"""Generates tests/golden/gl_ref_<variant>.npz: the targets a REAL OpenGL implementation (Mesa llvmpipe) produces when it runs the
reference's own shader text on the fixture scenes — oracle/_ref/glref, see oracle/glref/glref_harness.cpp and tests/test_gl_ref.py.
Run in the build container (needs /root/reference and the Mesa libGL):   python tests/golden/make_gl_golden.py
The fixtures travel; tests/test_gl_golden.py (CPU: the oracle, -m gpu: the CUDA path through the C ABI) compares against them on
machines that have neither the reference tree nor a GL stack.

Scenes: tests/fixtures.py:single_level_copy(gl_scene_of(name)) — level-0 texture filtering — rendered with float texel storage
(GLREF_FLOAT_TEXTURES), i.e. with llvmpipe's two sampling shortcuts out of the way, so that the colour target is comparable too.
Stored per variant: class / instance index (u16), vertex ids (u32 x 3), object coordinates + depth (f32 x 4), camera-space normals
(f16 x 4), RGBA8 colour.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import glref_util  # noqa: E402
import fixtures  # noqa: E402

VARIANTS = ["tabletop", "three_lights", "near_clip", "alpha_test", "pbr_textures", "low_poly_closeup", "sticker", "projective"]


def main():
    why = glref_util.available()
    if why:
        raise SystemExit(why)
    for name in VARIANTS:
        sc = fixtures.single_level_copy(fixtures.gl_scene_of(name))
        g = glref_util.render(sc, env={"GLREF_FLOAT_TEXTURES": "1"})
        path = os.path.join(HERE, f"gl_ref_{name}.npz")
        np.savez_compressed(path, class_index=g["class_index"], instance_index=g["instance_index"], vertex_index=g["vertex_index"][..., :3].copy(),
                            coord=g["coord"], normals=g["normals"].astype(np.float16), rgb=g["rgb"])
        print(name, os.path.getsize(path) // 1024, "KiB")
    post_scenes()
    bench_frame()


def post_scenes():
    """'ssao' and 'ibl' (SSAO + sky box + image-based lighting): the frame plus, for 'ibl', the light maps GL's run of LightMap::load's
    passes produced at 64 / 16 / 32 / 64 — the tests render with exactly those maps (LightMapData.maps), so the fixtures pin the
    render path; the precompute is compared in tests/test_gl_ref.py."""
    for name in ("ssao", "ibl"):
        sc = fixtures.gl_post_scene(name)
        g = glref_util.render(sc, env={"GLREF_FLOAT_TEXTURES": "1"}, lightmap_sizes=(64, 16, 32, 64))
        extra = {}
        if "lightmap" in g:
            extra = dict(zip(("lm_env0", "lm_irr", "lm_pre", "lm_lut"), g["lightmap"]))
        path = os.path.join(HERE, f"gl_ref_post_{name}.npz")
        np.savez_compressed(path, class_index=g["class_index"], instance_index=g["instance_index"], vertex_index=g["vertex_index"][..., :3].copy(),
                            coord=g["coord"], normals=g["normals"].astype(np.float16), rgb=g["rgb"], hdr=g["hdr"].astype(np.float16), **extra)
        print("post", name, os.path.getsize(path) // 1024, "KiB")


def bench_frame():
    """Scene 1 of the headline workload (bench.py C3, 640x480, 20 objects, 328 k triangles): ids, depth and colour only (file size)."""
    import bench
    sc = bench.build_scenes("C3", bench.build_pool(), None, 1, 2)[0]
    g = glref_util.render(sc, env={"GLREF_FLOAT_TEXTURES": "1"})
    path = os.path.join(HERE, "gl_ref_bench_c3.npz")
    np.savez_compressed(path, class_index=g["class_index"], instance_index=g["instance_index"], vertex_index=g["vertex_index"][..., :3].copy(),
                        depth=g["coord"][..., 3].copy(), rgb=g["rgb"])
    print("bench_c3", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
