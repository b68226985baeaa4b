"""Against the committed OpenGL goldens (tests/golden/gl_ref_*.npz, made by tests/golden/make_gl_golden.py): the targets Mesa llvmpipe
produced running the reference's own shader text — a real OpenGL implementation, independent of this repository's raster contract.
The CPU test holds the oracle to them, the -m gpu test holds the CUDA path (through the C ABI) to them DIRECTLY. Neither needs the
reference tree or a GL stack at run time. Thresholds follow tests/test_gl_ref.py (measured there: 0 - 11 id pixels per 76 800)."""
import os

import numpy as np
import pytest

import fixtures
from stillleben_b200 import abi

GL_GOLDEN = ["tabletop", "three_lights", "near_clip", "alpha_test", "pbr_textures", "low_poly_closeup", "sticker", "projective"]


def check_against_gl(out, name, prefix="gl_ref_", rgb_budget=200):
    z = np.load(os.path.join(fixtures.GOLDEN, f"{prefix}{name}.npz"))
    bad = (((z["coord"][..., 3] == abi.INVALID_COORD) != (out["coord"][..., 3] == abi.INVALID_COORD))
           | np.any(z["instance_index"] != out["instance_index"], axis=-1) | np.any(z["class_index"] != out["class_index"], axis=-1)
           | np.any(z["vertex_index"] != out["vertex_index"][..., :3], axis=-1))
    n_bad = int(bad.sum())
    assert n_bad <= 24, (name, n_bad)                       # visibility: ids + coverage of GL's rasteriser
    ok = ~bad
    d = np.abs(z["coord"] - out["coord"])[ok]
    assert d.max() <= 2e-3, (name, float(d.max()))          # object coordinates and depth
    n = np.abs(z["normals"].astype(np.float32) - out["normals"]).max(-1)[ok]
    assert int((n > 1e-2).sum()) <= 40, (name, int((n > 1e-2).sum()))
    d8 = np.abs(z["rgb"].astype(int) - out["rgb"].astype(int)).max(-1)[ok]
    assert int((d8 > 1).sum()) <= rgb_budget, (name, int((d8 > 1).sum()))   # colour through PBR + GL's own PCF compare + tone map
    return n_bad, int((d8 > 1).sum())


@pytest.mark.parametrize("name", GL_GOLDEN)
def test_oracle_matches_opengl_golden(name):
    import oracle_util as ou
    sc = fixtures.single_level_copy(fixtures.gl_scene_of(name))
    check_against_gl(ou.render(sc), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GL_GOLDEN)
def test_cuda_matches_opengl_golden(gpu_ctx, name):
    sc = fixtures.single_level_copy(fixtures.gl_scene_of(name))
    res = gpu_ctx.render([sc], target_mask=abi.TARGETS_ALL)
    gpu_ctx.synchronize()
    check_against_gl(res.frame_dict(0), name)


def post_scene(name):
    """fixtures.gl_post_scene + for 'ibl' the light maps of the golden itself (GL's own precompute) as LightMapData.maps."""
    sc = fixtures.gl_post_scene(name)
    if sc.light_map is not None:
        z = np.load(os.path.join(fixtures.GOLDEN, f"gl_ref_post_{name}.npz"))
        sc.light_map.maps = (z["lm_env0"], z["lm_irr"], z["lm_pre"], z["lm_lut"])
    return sc


@pytest.mark.parametrize("name", ["ssao", "ibl"])
def test_oracle_post_passes_match_opengl_golden(name):
    """SSAO + bilateral apply; sky box + image-based lighting + SSAO (measured in tests/test_gl_ref.py: lit IBL pixels beyond one RGBA8
    level are the implicit-LOD lookups of the irradiance map / LUT, a few hundred of 76 800)."""
    import oracle_util as ou
    sc = post_scene(name)
    assets = ou.OracleAssets()
    if sc.light_map is not None:
        assets.set_lightmap_maps(sc.light_map, *sc.light_map.maps)
    check_against_gl(ou.render(sc, assets), name, prefix="gl_ref_post_", rgb_budget=60 if name == "ssao" else 900)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ssao", "ibl"])
def test_cuda_post_passes_match_opengl_golden(gpu_ctx, name):
    sc = post_scene(name)
    res = gpu_ctx.render([sc], target_mask=abi.TARGETS_ALL)
    gpu_ctx.synchronize()
    check_against_gl(res.frame_dict(0), name, prefix="gl_ref_post_", rgb_budget=60 if name == "ssao" else 900)


def check_bench_frame(out):
    """The headline workload's own frame (bench.py C3, scene 1) against GL: ids / coverage, depth, colour (mip-mapped 8-bit textures on
    the CUDA / oracle side, float texels on GL's: the colour budget is the LOD approximation's, tests/test_gl_ref.py)."""
    z = np.load(os.path.join(fixtures.GOLDEN, "gl_ref_bench_c3.npz"))
    bad = (((z["depth"] == abi.INVALID_COORD) != (out["coord"][..., 3] == abi.INVALID_COORD))
           | np.any(z["instance_index"] != out["instance_index"], axis=-1) | np.any(z["class_index"] != out["class_index"], axis=-1)
           | np.any(z["vertex_index"] != out["vertex_index"][..., :3], axis=-1))
    assert int(bad.sum()) <= 40, int(bad.sum())                     # measured: 6 of 307 200
    assert np.abs(z["depth"] - out["coord"][..., 3])[~bad].max() <= 1e-3
    d8 = np.abs(z["rgb"].astype(int) - out["rgb"].astype(int)).max(-1)[~bad]
    assert int((d8 > 8).sum()) <= 3000, int((d8 > 8).sum())


def bench_scene():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    return bench.build_scenes("C3", bench.build_pool(), None, 1, 2)[0]


def test_oracle_bench_frame_matches_opengl_golden():
    import oracle_util as ou
    check_bench_frame(ou.render(bench_scene()))


@pytest.mark.gpu
def test_cuda_bench_frame_matches_opengl_golden(gpu_ctx):
    res = gpu_ctx.render([bench_scene()], target_mask=abi.TARGETS_ALL)
    gpu_ctx.synchronize()
    check_bench_frame(res.frame_dict(0))
