"""Size-independent properties at BASELINE.json's full sizes (640x480, 20 objects x 16k triangles),
where running the oracle on every frame would take too long: determinism / idempotence, batch
independence, invariants of the targets, plus oracle parity on a sample of the frames."""
import hashlib

import numpy as np
import pytest

import oracle_util as ou
import parity
from stillleben_b200 import abi, synth

pytestmark = pytest.mark.gpu

N = 24


@pytest.fixture(scope="module")
def batch(gpu_ctx):
    pool = synth.mesh_pool(21)
    scenes = [synth.tabletop_scene(pool, 1000 + s) for s in range(N)]
    res = gpu_ctx.render(scenes, target_mask=abi.TARGETS_ALL)
    gpu_ctx.synchronize()
    return pool, scenes, res


def digest(res, i):
    h = hashlib.sha256()
    for t in range(abi.NUM_TARGETS):
        h.update(res.numpy(t, i, 1).tobytes())
    return h.hexdigest()


def test_idempotent_and_order_independent(gpu_ctx, batch):
    pool, scenes, res = batch
    again = gpu_ctx.render(scenes[::-1], target_mask=abi.TARGETS_ALL)      # reversed order, fresh result
    gpu_ctx.synchronize()
    d0 = [digest(res, i) for i in range(N)]
    d1 = [digest(again, N - 1 - i) for i in range(N)]
    assert d0 == d1
    assert len(set(d0)) == N                     # every scene is different
    # checksum of checksums is stable across a third render into the SAME result (buffers reused)
    gpu_ctx.render(scenes, result=res)
    gpu_ctx.synchronize()
    assert hashlib.sha256("".join(d0).encode()).hexdigest() == hashlib.sha256("".join(digest(res, i) for i in range(N)).encode()).hexdigest()


def test_target_invariants(batch):
    pool, scenes, res = batch
    inst = res.numpy(abi.TARGET_INSTANCE)[..., 0]
    cls = res.numpy(abi.TARGET_CLASS)[..., 0]
    coord = res.numpy(abi.TARGET_COORD)
    vid = res.numpy(abi.TARGET_VERTEX_INDEX)
    bary = res.numpy(abi.TARGET_BARY)
    nrm = res.numpy(abi.TARGET_NORMAL)
    cam = res.numpy(abi.TARGET_CAM_COORD)
    rgb = res.numpy(abi.TARGET_RGB)
    assert inst.max() <= 20 and cls.max() <= 21
    assert ((inst == 0) == (cls == 0)).all()
    covered = coord[..., 3] < 2999.0
    depth = coord[..., 3][covered]
    assert depth.min() >= 0.1 - 1e-4 and depth.max() <= 10.0 + 1e-3        # near / far planes (scene.cpp:227-228)
    assert (coord[~covered] == 3000.0).all() and (cam[~covered] == 3000.0).all()
    assert (rgb[..., 3][covered] == 255).all() and (rgb[..., 3][~covered] == 0).all()
    obj = inst > 0
    assert (vid[..., 0][obj] > 0).all() and (vid[..., 3] == 0).all()
    assert (vid[..., :3][obj] <= 8385).all()
    np.testing.assert_allclose(bary[..., :3][obj].sum(-1), 1.0, rtol=2e-5)
    np.testing.assert_allclose(np.linalg.norm(nrm[..., :3][covered], axis=-1), 1.0, rtol=1e-5)
    np.testing.assert_allclose(cam[..., 2][covered], coord[..., 3][covered], rtol=1e-6)   # coord.w is camera z
    # object coordinates lie inside the (pretransformed) bounding sphere of the instance's mesh
    for f in range(0, N, 6):
        sc = scenes[f]
        for k, o in enumerate(sc.objects):
            m = inst[f] == k + 1
            if not m.any():
                continue
            r = 0.5 * np.linalg.norm((o.mesh.bbox_max - o.mesh.bbox_min) * o.pretransform[0, 0])
            assert np.linalg.norm(coord[f][m][:, :3], axis=-1).max() <= r * 1.001 + 1e-5


def test_sampled_frames_match_oracle(batch):
    pool, scenes, res = batch
    assets = ou.OracleAssets()
    for i in (0, 7, 19):
        ref = ou.render(scenes[i], assets, want_hdr=False)
        parity.assert_parity(res.frame_dict(i), ref, rgb_outliers=4)


def test_raster_paths_agree_bit_exactly(gpu_ctx, batch):
    """The visibility result is the per-pixel minimum key whichever way a triangle is rasterised: all tiled
    (0/0), all small ones by one thread (4096/4096), all by one warp (0/4096), the default mix — every
    target of every frame must come out byte-identical (SLB_OPT_DIRECT_MAX / SLB_OPT_WARP_MAX)."""
    pool, scenes, _ = batch
    try:
        # the per-frame HugeShade records change the ROUNDING of the float targets of huge-triangle pixels (not the ids): this
        # test is about visibility, so both sides shade through the generic path (the records have their own test below)
        gpu_ctx.set_option(abi.OPT_HUGE_PREPARE, 0)
        res = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL)
        gpu_ctx.synchronize()
        ref = [digest(res, i) for i in range(6)]
        for direct_max, warp_max, huge in ((0, 0, 0), (4096, 4096, 1), (0, 4096, 0), (16, 256, 1), (128, 4096, 0), (0, 0, 1)):
            gpu_ctx.set_option(abi.OPT_DIRECT_MAX, direct_max)
            gpu_ctx.set_option(abi.OPT_WARP_MAX, warp_max)
            gpu_ctx.set_option(abi.OPT_HUGE_IN_SHADE, huge)          # huge triangles per pixel in the shade kernel / tile-binned
            again = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL)
            gpu_ctx.synchronize()
            assert [digest(again, i) for i in range(6)] == ref, (direct_max, warp_max, huge)
    finally:
        gpu_ctx.set_option(abi.OPT_DIRECT_MAX, 128)
        gpu_ctx.set_option(abi.OPT_WARP_MAX, 4096)
        gpu_ctx.set_option(abi.OPT_HUGE_IN_SHADE, 1)
        gpu_ctx.set_option(abi.OPT_HUGE_PREPARE, 1)


def test_shadow_block_masks_change_nothing(gpu_ctx, batch):
    """SLB_OPT_SHADOW_MASK: PCF footprints over untouched 8x8 texel blocks skip their 25 taps — every target of every frame
    must stay byte-identical, with one and with three shadow lights, whichever raster path fills the maps."""
    import fixtures
    pool, scenes, _ = batch
    extra = [fixtures.variant("three_lights"), fixtures.variant("low_poly_closeup"), fixtures.variant("near_clip")]
    # a sparse scene at full HD: few shadow triangles per set-up block, i.e. warps with a handful of active lanes whose joint
    # box spans several mask rows (the case a lane-strided marking loop gets wrong)
    from stillleben_b200 import synth
    extra.append(synth.tabletop_scene(fixtures.small_pool(), 31, n_objects=8, width=1920, height=1080, intrinsics=None, n_lights=3))
    extra.append(fixtures.variant("c2_shape"))
    try:
        gpu_ctx.set_option(abi.OPT_HUGE_PREPARE, 0)      # which triangles are "huge" depends on the raster thresholds: keep one shading path
        on6 = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL)
        def extras():
            out = []
            for sc in extra:           # different viewports: one call each
                r = gpu_ctx.render([sc], target_mask=abi.TARGETS_ALL)
                gpu_ctx.synchronize()
                out.append(digest(r, 0))
            return out
        gpu_ctx.set_option(abi.OPT_SHADOW_MASK, 0)
        on6 = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL)
        gpu_ctx.synchronize()
        on = [digest(on6, i) for i in range(6)] + extras()
        for direct_max, warp_max in ((128, 4096), (0, 0)):
            gpu_ctx.set_option(abi.OPT_DIRECT_MAX, direct_max)
            gpu_ctx.set_option(abi.OPT_WARP_MAX, warp_max)
            for mask in (0, 1):
                gpu_ctx.set_option(abi.OPT_SHADOW_MASK, mask)
                a = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL)
                gpu_ctx.synchronize()
                assert [digest(a, i) for i in range(6)] + extras() == on, (direct_max, warp_max, mask)
    finally:
        gpu_ctx.set_option(abi.OPT_DIRECT_MAX, 128)
        gpu_ctx.set_option(abi.OPT_WARP_MAX, 4096)
        gpu_ctx.set_option(abi.OPT_SHADOW_MASK, 1)
        gpu_ctx.set_option(abi.OPT_HUGE_PREPARE, 1)


def test_huge_shade_records_agree_with_the_generic_path(gpu_ctx, batch):
    """SLB_OPT_HUGE_PREPARE: huge sub-triangles (the background plane, close-up faces) shaded from per-frame records of the
    vertex stage's outputs instead of a per-pixel re-set-up — ids identical, float targets within rounding, colour within
    1 LSB; incl. clipped planes, textured planes and triangles that fall back to the generic path."""
    import fixtures
    scenes = list(batch[1][:3]) + [fixtures.variant(n) for n in ("near_clip", "plane_texture", "low_poly_closeup", "projective", "sticker")]
    groups = [scenes[:3], scenes[3:]]
    outs = {}
    try:
        for mode in (1, 0):
            gpu_ctx.set_option(abi.OPT_HUGE_PREPARE, mode)
            outs[mode] = [gpu_ctx.render(g, target_mask=abi.TARGETS_ALL) for g in groups]
            gpu_ctx.synchronize()
    finally:
        gpu_ctx.set_option(abi.OPT_HUGE_PREPARE, 1)
    for gi, g in enumerate(groups):
        for i in range(len(g)):
            a, b = outs[1][gi].frame_dict(i), outs[0][gi].frame_dict(i)
            st = parity.compare(a, b)
            for name, s_ in st.items():
                if name in parity.EXACT:
                    assert s_["mismatch"] == 0, (gi, i, name, s_)
                elif name == "rgb":
                    assert s_["over1"] <= 2, (gi, i, s_)
                else:
                    assert s_["bad"] == 0, (gi, i, name, s_)


def test_lean_and_full_shade_kernels_agree(gpu_ctx, batch):
    """Sub-batches without material textures / stickers / light maps run a lean instantiation of the shade kernel
    (SLB_OPT_LEAN_SHADE): ids identical, float targets equal up to contraction differences, colour within 1 LSB."""
    pool, scenes, res = batch
    try:
        gpu_ctx.set_option(abi.OPT_LEAN_SHADE, 0)
        full = gpu_ctx.render(scenes[:4], target_mask=abi.TARGETS_ALL)
        gpu_ctx.synchronize()
    finally:
        gpu_ctx.set_option(abi.OPT_LEAN_SHADE, 1)
    for i in range(4):
        a, b = res.frame_dict(i), full.frame_dict(i)
        st = parity.compare(a, b)
        for name, s_ in st.items():
            if name in parity.EXACT:
                assert s_["mismatch"] == 0, (name, s_)
            elif name == "rgb":
                assert s_["over1"] == 0, s_
            else:
                assert s_["bad"] == 0 and s_["max_abs"] < 1e-4 * 3000, (name, s_)


def test_shadow_map_generations_wrap(gpu_ctx):
    """Shadow-map texels carry an 8-bit generation tag instead of being cleared for every sub-batch
    (DFrame::shadow_tagbits); after 255 sub-batches the pool is really cleared and the count restarts. Rendering
    the same scenes through > 2 x 255 single-scene sub-batches must keep producing identical frames."""
    import fixtures
    a, b = fixtures.small_tabletop_scene(), fixtures.variant("three_lights", 160, 120)
    gpu_ctx.set_option(abi.OPT_MAX_SUBBATCH, 1)
    try:
        ref = gpu_ctx.render([a, b], target_mask=abi.TARGETS_SIX)
        gpu_ctx.synchronize()
        want = [hashlib.sha256(ref.numpy(abi.TARGET_RGB, i, 1).tobytes() + ref.numpy(abi.TARGET_COORD, i, 1).tobytes()).hexdigest() for i in range(2)]
        for it in range(270):                       # 540 sub-batches with 1 or 3 shadow maps each
            res = gpu_ctx.render([a, b], result=ref)
            if it % 30 == 0 or it > 250:
                gpu_ctx.synchronize()
                got = [hashlib.sha256(res.numpy(abi.TARGET_RGB, i, 1).tobytes() + res.numpy(abi.TARGET_COORD, i, 1).tobytes()).hexdigest() for i in range(2)]
                assert got == want, it
    finally:
        gpu_ctx.set_option(abi.OPT_MAX_SUBBATCH, 64)


def test_two_stream_pipeline_changes_nothing(gpu_ctx, batch):
    """SLB_OPT_OVERLAP: first phases of the sub-batches on the auxiliary stream, second phases on the render stream. Every target of every
    frame must stay byte-identical — with sub-batches of two frames (six frames = three sub-batches: both scratch sets are reused inside
    one call), over repeated calls (the reuse is also ordered ACROSS calls), and with work queued on a caller's stream around it."""
    import torch
    pool, scenes, _ = batch
    try:
        gpu_ctx.set_option(abi.OPT_MAX_SUBBATCH, 2)
        gpu_ctx.set_option(abi.OPT_OVERLAP, 0)
        res = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL)
        gpu_ctx.synchronize()
        ref = [digest(res, i) for i in range(6)]
        gpu_ctx.set_option(abi.OPT_OVERLAP, 1)
        stream = torch.cuda.Stream()
        for attempt in range(4):
            if attempt % 2:      # on a caller's stream, back to back with another render whose result is read after both
                with torch.cuda.stream(stream):
                    other = gpu_ctx.render(scenes[3:6] + scenes[:3], target_mask=abi.TARGETS_ALL, stream=stream.cuda_stream)
                    again = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL, stream=stream.cuda_stream)
                stream.synchronize()
                gpu_ctx.synchronize()
                assert [digest(other, i) for i in range(6)] == ref[3:] + ref[:3], attempt
            else:
                again = gpu_ctx.render(scenes[:6], target_mask=abi.TARGETS_ALL)
                gpu_ctx.synchronize()
            assert [digest(again, i) for i in range(6)] == ref, attempt
    finally:
        gpu_ctx.set_option(abi.OPT_MAX_SUBBATCH, 64)
        gpu_ctx.set_option(abi.OPT_OVERLAP, 1)
