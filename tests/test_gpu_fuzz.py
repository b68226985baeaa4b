"""The CUDA path against the oracle on the seeded corner-case scenes of fixtures.fuzz_scene — the same forty scenes tests/test_gl_ref.py
runs through a real OpenGL implementation (viewports from 1x1 up, geometry crossing the near plane / behind the camera, 100 - 150 degree
fields of view, sub-pixel triangles, grazing views, hidden objects, non-casters, 0 - 3 shadow lights). Same bar as tests/test_gpu_parity.py:
ids bit-exact, float targets 1e-3, RGBA8 one level."""
import pytest

import fixtures
import oracle_util as ou
import parity
from stillleben_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("block", range(4))
def test_fuzzed_corner_cases_match_oracle(gpu_ctx, block):
    gpu_ctx.set_option(abi.OPT_KEEP_HDR, 1)
    try:
        for seed in range(10 * block, 10 * block + 10):
            sc = fixtures.fuzz_scene(seed)
            res = gpu_ctx.render([sc], target_mask=abi.TARGETS_ALL)
            gpu_ctx.synchronize()
            out = res.frame_dict(0)
            out["hdr"] = res.hdr(0)
            try:
                parity.assert_parity(out, ou.render(sc), rgb_outliers=4, hdr_outlier_frac=2e-4)
            except AssertionError as e:
                raise AssertionError(f"seed {seed}: {e}") from None
    finally:
        gpu_ctx.set_option(abi.OPT_KEEP_HDR, 0)
