"""Host-side logic: camera matrices, descriptor marshalling, sharding and the asset arena."""
import ctypes as C

import numpy as np
import pytest

from stillleben_b200 import abi, desc, dist, synth


def test_intrinsics_projection_matches_closed_form():
    # SURVEY Appendix A.2 (src/scene.cpp:222-253)
    fx, fy, cx, cy, W, H = 1066.778, 1067.487, 312.9869, 241.3109, 640, 480
    P = desc.intrinsics_projection(fx, fy, cx, cy, W, H)
    p = np.array([0.1, -0.05, 1.3, 1.0], np.float64)
    c = P.astype(np.float64) @ p
    ndc = c[:3] / c[3]
    assert abs((ndc[0] * 0.5 + 0.5) * W - (fx * p[0] / p[2] + cx)) < 1e-3
    assert abs((ndc[1] * 0.5 + 0.5) * H - (fy * p[1] / p[2] + cy)) < 1e-3
    n, f = 0.1, 10.0
    assert abs(ndc[2] - ((f + n) / (f - n) - 2 * f * n / ((f - n) * p[2]))) < 1e-5
    assert c[3] == p[2]


def test_look_at_is_rigid_and_points_forward():
    m = desc.look_at_pose((4, 0, 0), (0, 0, 0))
    np.testing.assert_allclose(m[:3, :3].T @ m[:3, :3], np.eye(3), atol=1e-6)
    np.testing.assert_allclose(m[:3, 2], [-1, 0, 0], atol=1e-6)       # +z forward
    inv = desc.inverted_rigid(m)
    np.testing.assert_allclose(inv @ m, np.eye(4), atol=1e-6)


def test_desc_batch_marshalling():
    pool = synth.mesh_pool(2, nu=8, nv=4, tex_size=8)
    sc = synth.tabletop_scene(pool, 1, n_objects=3)
    sc.objects[1].instance_index = 0        # auto -> position in scene, 1-based (scene.cpp:285-287)
    handles = {}
    b = desc.DescBatch([sc], lambda o: handles.setdefault(id(o), 1000 + len(handles)))
    d = b.scenes[0]
    assert (d.width, d.height, d.n_objects) == (640, 480, 3)
    assert d.objects[1].instance_index == 2
    assert d.objects[0].mesh == handles[id(sc.objects[0].mesh)]
    # column-major storage of a row-major numpy matrix
    assert abs(d.projection[11] - sc.projection[3, 2]) < 1e-7 and d.projection[11] == 1.0
    assert d.background_plane_size[0] == 3.0 and d.light_map is None


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1024, 1025):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                lo, hi = dist.shard_range(n, r, w)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_asset_arena_round_trip():
    pool = synth.mesh_pool(3, nu=16, nv=8, tex_size=16)
    header, arena = dist.pack_meshes(pool)
    back = dist.unpack_meshes(header, arena)
    assert len(back) == len(pool)
    for a, b in zip(pool, back):
        assert (a.vertices == b.vertices).all() and (a.indices == b.indices).all()
        assert a.submeshes == b.submeshes
        assert all((x.pixels == y.pixels).all() for x, y in zip(a.images, b.images))
        assert a.geometry_bytes() == b.geometry_bytes()


def test_vertex_dtype_is_reference_layout():
    # src/mesh_tools/consolidate.cpp:53-61
    dt = abi.VERTEX_DTYPE
    assert dt.itemsize == 68
    assert [dt.fields[k][1] for k in ("position", "uv", "color", "tangent", "vertex_index", "normal")] == [0, 12, 20, 36, 52, 56]


def test_cpulist_parser_and_numa_binding_are_harmless_without_a_gpu():
    from stillleben_b200 import dist
    assert dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert dist._parse_cpulist("") == set()
    import os
    before = os.sched_getaffinity(0)
    assert dist.bind_to_gpu_numa_node(0) in (None, 0, 1, 2, 3, 4, 5, 6, 7)     # no GPU here: None, and nothing changes
    if dist.bind_to_gpu_numa_node(0) is None:
        assert os.sched_getaffinity(0) == before


def test_oracle_recompute_normals_matches_float64_restating():
    """orc_mesh_recompute_normals (the twin of Mesh::recomputeNormals, src/mesh.cpp:763-816) against a float64 numpy
    restatement: area-weighted face normals summed per vertex; the position update adds to one-based ids."""
    import ctypes as C
    import oracle_util as ou
    from stillleben_b200 import abi, synth
    mesh = synth.shape_mesh("blob", 11, nu=16, nv=8, textured=False)
    assets = ou.OracleAssets()
    h = assets.handle_of(mesh)
    L = ou.lib()
    L.orc_mesh_update_positions_and_colors.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.orc_mesh_read_vertices.argtypes = [C.c_void_p, C.c_void_p]
    rng = np.random.RandomState(0)
    n = len(mesh.vertices)
    ids = np.arange(1, n + 1, dtype=np.int32)[::3].copy()
    dpos = (rng.normal(size=(len(ids), 3)) * 0.03).astype(np.float32)
    assert L.orc_mesh_update_positions_and_colors(h, ids.ctypes.data, len(ids), dpos.ctypes.data, None) == 0
    out = np.empty(n, abi.VERTEX_DTYPE)
    L.orc_mesh_read_vertices(h, out.ctypes.data)
    pos = mesh.vertices["position"].astype(np.float64).copy()
    pos[ids - 1] += dpos
    np.testing.assert_allclose(out["position"], pos, rtol=0, atol=1e-7)
    tri = mesh.indices.reshape(-1, 3)
    fn = np.cross(pos[tri[:, 0]] - pos[tri[:, 1]], pos[tri[:, 0]] - pos[tri[:, 2]])        # normal * area = the cross product itself
    acc = np.zeros_like(pos)
    for k in range(3):
        np.add.at(acc, tri[:, k], fn)
    ok = np.linalg.norm(acc, axis=1) > 1e-12
    expect = acc[ok] / np.linalg.norm(acc[ok], axis=1, keepdims=True)
    np.testing.assert_allclose(out["normal"][ok], expect, atol=2e-5)
    assert L.orc_mesh_update_positions_and_colors(h, np.array([0], np.int32).ctypes.data, 1, dpos.ctypes.data, None) == -1


def test_quaternion_helpers_round_trip():
    """py_magnum.cpp:83-113: [x y z w] <-> 3x3, every branch of Quaternion::fromMatrix."""
    import torch
    from stillleben_b200 import sl
    rng = np.random.RandomState(3)
    qs = [rng.normal(size=4) for _ in range(50)] + [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.7, 0.7, 0.01, 0.01]]
    for q in qs:
        q = np.asarray(q, np.float64) / np.linalg.norm(q)
        R = sl.quat_to_matrix(torch.tensor(q, dtype=torch.float32))
        assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-5) and abs(float(torch.det(R)) - 1) < 1e-5
        q2 = sl.matrix_to_quat(R).numpy().astype(np.float64)
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-5
    with pytest.raises(ValueError):
        sl.quat_to_matrix(torch.zeros(3))


def test_obj_front_end(tmp_path):
    """objfile.load: (v, vt, vn) triples joined, quads fanned, negative indices, MTL colour + texture, generated normals."""
    from PIL import Image
    from stillleben_b200 import objfile
    Image.fromarray(np.arange(4 * 4 * 3, dtype=np.uint8).reshape(4, 4, 3)).save(tmp_path / "t.png")
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 0.5 0.25 1\nmap_Kd t.png\nnewmtl b\nKd 1 0 0\nd 0.5\n")
    (tmp_path / "m.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                                    "usemtl a\nf 1/1/1 2/2/1 3/3/1 4/4/1\nusemtl b\nv 0 0 1\nv 1 0 1\nv 0 1 1\nf -3 -2 -1\n")
    m = objfile.load(str(tmp_path / "m.obj"))
    assert len(m.vertices) == 7 and m.submeshes == [(0, 6, 0), (6, 3, 1)]
    assert m.indices.tolist() == [0, 1, 2, 0, 2, 3, 4, 5, 6]
    assert m.vertices["vertex_index"].tolist() == list(range(1, 8))
    assert np.allclose(m.vertices["normal"][:4], [0, 0, 1]) and np.allclose(m.vertices["normal"][4:], [0, 0, 1])     # generated
    assert np.allclose(m.vertices["tangent"][:4], [1, 0, 0, 1]) and np.allclose(m.vertices["uv"][2], [1, 1])
    assert m.materials[0].base_color == (0.5, 0.25, 1.0, 1.0) and m.materials[0].tex_base_color == 0
    assert m.materials[1].base_color == (1.0, 0.0, 0.0, 0.5) and m.materials[1].tex_base_color == -1
    assert m.images[0].pixels[0, 0].tolist() == [36, 37, 38]            # row 0 = bottom row of the picture


def test_weighted_sharding():
    """dist.shard_sizes_weighted / shard_range_weighted: sizes proportional to the ranks' link bandwidths, exact total,
    contiguous and disjoint ranges; equal weights reproduce shard_range."""
    w8 = [11.58, 11.59, 11.59, 11.61, 18.15, 18.2, 18.23, 18.26]          # profiles/r02_d2h_probe_n8.json
    sizes = dist.shard_sizes_weighted(1024, w8)
    assert sum(sizes) == 1024 and all(abs(s - 1024 * w / sum(w8)) < 1.0 for s, w in zip(sizes, w8))
    # equal time per rank: bytes over bandwidth within one scene of each other
    t = [s / w for s, w in zip(sizes, w8)]
    assert max(t) - min(t) <= 1.0 / min(w8) + 1e-9
    ranges = [dist.shard_range_weighted(1024, r, w8) for r in range(8)]
    assert ranges[0][0] == 0 and ranges[-1][1] == 1024 and all(ranges[r][1] == ranges[r + 1][0] for r in range(7))
    for n, world in ((1024, 8), (10, 4), (3, 8), (0, 2)):
        assert [dist.shard_range_weighted(n, r, [1.0] * world) for r in range(world)] != [] and \
            sum(dist.shard_sizes_weighted(n, [1.0] * world)) == n
        assert sorted(dist.shard_sizes_weighted(n, [1.0] * world), reverse=True) == sorted((dist.shard_range(n, r, world)[1] - dist.shard_range(n, r, world)[0]
                                                                                            for r in range(world)), reverse=True)
    assert dist.shard_sizes_weighted(7, [0.0, 0.0]) in ([4, 3], [3, 4])                                # degenerate weights: equal shards
    assert dist.shard_sizes_weighted(5, [1.0, 0.0, 3.0]) == [1, 0, 4]


def test_neg_iou_loss_matches_its_definition():
    import torch
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("slb_losses", os.path.join(os.path.dirname(os.path.dirname(__file__)), "stillleben", "losses.py"))
    losses = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(losses)
    g = torch.Generator().manual_seed(0)
    p, t = torch.rand(3, 1, 8, 8, generator=g), (torch.rand(3, 1, 8, 8, generator=g) > 0.5).float()
    loss, img = losses.neg_iou_loss(p, t)
    inter, union = (p * t).sum((1, 2, 3)), (p + t - p * t).sum((1, 2, 3)) + 1e-6
    assert abs(float(loss) - float(1 - (inter / union).mean())) < 1e-6
    assert img.shape == p.shape and not img.requires_grad
    assert float(losses.neg_iou_loss(t, t)[0]) < 1e-5


def test_profiling_timer_context_and_decorator(capsys):
    """stillleben.profiling.Timer as tests/test_python.py uses it: decorator + nested context manager, silent unless enabled."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("slb_profiling", os.path.join(os.path.dirname(os.path.dirname(__file__)), "stillleben", "profiling.py"))
    prof = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(prof)

    @prof.Timer("stage")
    def stage():
        return 7

    with prof.Timer("quiet"):
        assert stage() == 7
    assert capsys.readouterr().out == ""
    prof.Timer.enabled = True
    try:
        with prof.Timer("outer") as t:
            stage()
            with prof.Timer("inner"):
                stage()
        lines = capsys.readouterr().out.splitlines()
        assert lines[0] == "Timings:" and [ln.split()[0] for ln in lines[1:]] == ["outer", "stage", "inner", "stage"]
        assert lines[2].startswith("  stage") and lines[4].startswith("    stage") and t.duration > 0
    finally:
        prof.Timer.enabled = False


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_ply_front_end(tmp_path, fmt):
    """plyfile.load: the three PLY encodings give the same mesh; normals / colours / uv read or generated; quads fanned;
    TextureFile comment -> base-colour texture (row 0 = bottom row)."""
    from PIL import Image
    from stillleben_b200 import plyfile
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1]], np.float32)
    faces = [(0, 1, 2, 3), (0, 1, 4), (1, 2, 4), (2, 3, 4), (3, 0, 4)]
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (5, 1))
    col = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [10, 20, 30], [255, 255, 255]], np.uint8)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.5]], np.float32)
    Image.fromarray(np.arange(4 * 4 * 3, dtype=np.uint8).reshape(4, 4, 3)).save(tmp_path / "t.png")
    plyfile._write_test_ply(tmp_path / "a.ply", pos, faces, fmt, normals=nrm, colors=col, uv=uv, texture_file="t.png")
    m = plyfile.load(str(tmp_path / "a.ply"))
    assert len(m.vertices) == 5 and m.indices.tolist() == [0, 1, 2, 0, 2, 3, 0, 1, 4, 1, 2, 4, 2, 3, 4, 3, 0, 4]
    np.testing.assert_allclose(m.vertices["position"], pos)
    np.testing.assert_allclose(m.vertices["normal"], nrm)
    np.testing.assert_allclose(m.vertices["color"][:, :3], col / 255.0, atol=1e-7)
    assert (m.vertices["color"][:, 3] == 1.0).all() and (m.vertices["tangent"][:, 3] == 1.0).all()
    np.testing.assert_allclose(m.vertices["uv"], uv)
    assert m.vertices["vertex_index"].tolist() == [1, 2, 3, 4, 5]
    assert m.materials[0].tex_base_color == 0 and m.images[0].pixels[0, 0].tolist() == [36, 37, 38]
    assert m.submeshes == [(0, 18, 0)] and np.allclose(m.bbox_max, [1, 1, 1])
    # geometry only: normals generated, no texture, no colours
    plyfile._write_test_ply(tmp_path / "b.ply", pos, faces[1:], fmt)
    g = plyfile.load(str(tmp_path / "b.ply"))
    assert np.allclose(np.linalg.norm(g.vertices["normal"], axis=1), 1.0, atol=1e-5) and g.materials[0].tex_base_color == -1 and not g.images
    assert g.vertices["normal"][4, 2] > 0.5                       # the apex normal points up
    with pytest.raises(RuntimeError, match="not a PLY"):
        (tmp_path / "c.ply").write_bytes(b"solid nothing")
        plyfile.load(str(tmp_path / "c.ply"))
