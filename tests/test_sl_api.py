"""The host mirror of the reference's Python surface (stillleben_b200.sl), exercised the way the reference's
own tests use `import stillleben as sl` (tests/test_python.py:26-68, tests/basic.cpp:108-261,375-453)."""
import numpy as np
import pytest
import torch

import fixtures
import oracle_util as ou
import parity
from stillleben_b200 import sl


def test_requires_init_like_the_reference():
    if sl._ctx is not None:
        pytest.skip("context already created in this process")
    with pytest.raises(RuntimeError, match="Call sl::init"):       # py_context.cpp:69-75
        sl.Scene((640, 480))
    with pytest.raises(RuntimeError):
        sl.RenderPass()


def test_host_side_mesh_and_object_logic(monkeypatch):
    monkeypatch.setattr(sl, "_ctx", object())                       # host logic only: no device call below
    mesh = sl.Mesh.from_data(fixtures.load_mesh("cube_glb_mesh"))
    assert mesh.class_index == 1
    with pytest.raises(ValueError):
        mesh.class_index = 70000                                    # mesh.cpp:1083-1089
    mesh.center_bbox()
    mesh.scale_to_bbox_diagonal(0.5)
    assert abs(mesh.bbox.diagonal - 0.5) < 1e-6                     # tests/basic.cpp:128-129
    np.testing.assert_allclose(mesh.bbox.center.numpy(), 0.0, atol=1e-6)
    m = torch.eye(4); m[0, 0] = 2.0
    with pytest.raises(ValueError, match="not uniform"):            # mesh.cpp:1050-1073
        mesh.pretransform = m
    m = torch.eye(4) * 3.0; m[3, 3] = 1.0; m[0, 3] = 6.0
    mesh.pretransform = m
    np.testing.assert_allclose(mesh.pretransform.numpy(), m.numpy(), atol=1e-6)
    obj = sl.Object(mesh)
    with pytest.raises(ValueError):
        obj.instance_index = 1 << 16                                # object.cpp:376-382
    assert mesh.points.shape == (24, 3) and mesh.faces.shape == (36,)
    scene = sl.Scene((640, 480))
    scene.add_object(obj)
    assert obj.instance_index == 1                                  # tests/basic.cpp:152
    assert scene.light_colors[0].tolist() == [300.0, 300.0, 300.0] and float(scene.light_directions.abs().sum()) == 0.0
    assert scene.manual_exposure < 0
    with pytest.raises(RuntimeError, match="physics is not available"):
        scene.simulate(0.002)
    with pytest.raises(ValueError, match="unknown shading"):       # py_render_pass.cpp:244
        sl.RenderPass("toon")


def test_tabletop_placement_and_serialization(monkeypatch):
    """Scene.serialize / deserialize (scene.cpp:761-868, tests/test_python.py:68-88) and the documented NON-physical stand-in
    for simulate_tabletop_scene (scene.cpp:612-759): poses above the table, bounding spheres disjoint, plane pose set."""
    import os
    monkeypatch.setattr(sl, "_ctx", object())
    paths = [os.path.join(os.path.dirname(__file__), "golden", "assets", f) for f in ("pbr_patch.glb", "kitchen_sink.glb")]
    scene = sl.Scene((640, 480))
    scene.set_camera_intrinsics(1066.778, 1067.487, 312.9869, 241.3109)
    scene.set_camera_look_at(torch.tensor([0.6, 0.1, 0.5]), torch.tensor([0.0, 0.0, 0.05]))
    scene.ambient_light = torch.tensor([0.1, 0.2, 0.3])
    scene.light_directions = torch.tensor([[0.1, 0.2, -0.9], [0, 0, 0], [0, 0, 0]])
    scene.manual_exposure = 1.5
    file_meshes = []
    for path in paths:                         # (objects of one file share their mesh: the reference's deserialize() loads each file once)
        mesh = sl.Mesh(path)
        mesh.center_bbox()
        mesh.scale_to_bbox_diagonal(0.1 + 0.05 * len(file_meshes))
        mesh.class_index = len(file_meshes) + 3
        file_meshes.append(mesh)
    for k in range(4):
        o = sl.Object(file_meshes[k % 2])
        o.metallic, o.roughness, o.casts_shadows = 0.25 * k, 0.9 - 0.2 * k, bool(k % 2)
        scene.add_object(o)
    with pytest.warns(UserWarning, match="non-physical"):
        scene.simulate_tabletop_scene()
    spheres = []
    for o in scene.objects:
        c = o.pose() @ torch.cat([o.mesh.bbox.center, torch.ones(1)])
        r = o.mesh.bbox.diagonal / 2
        assert abs(float(c[2]) - (0.04 + r)) < 1e-5                      # resting on the table top
        R = o.pose()[:3, :3]
        assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-5)
        spheres.append((c[:3], r))
    for i in range(4):
        for j in range(i):
            assert float(torch.linalg.norm(spheres[i][0] - spheres[j][0])) >= spheres[i][1] + spheres[j][1] - 1e-5
    assert abs(float(scene.background_plane_pose[2, 3]) - 0.04) < 1e-7

    text = scene.serialize()
    assert "[object/mesh]" in text and text.count("[object]") == 4 and text.count("[light]") == 3
    scene2 = sl.Scene((320, 240))
    scene2.deserialize(text)
    assert scene2.viewport == (640, 480)
    np.testing.assert_allclose(scene2.projection_matrix().numpy(), scene.projection_matrix().numpy(), rtol=1e-6)
    np.testing.assert_allclose(scene2.camera_pose().numpy(), scene.camera_pose().numpy(), atol=1e-6)
    np.testing.assert_allclose(scene2.light_directions.numpy(), scene.light_directions.numpy(), atol=1e-7)
    np.testing.assert_allclose(scene2.background_plane_pose.numpy(), scene.background_plane_pose.numpy(), atol=1e-7)
    assert scene2.manual_exposure == 1.5 and np.allclose(scene2.ambient_light.numpy(), [0.1, 0.2, 0.3])
    for a, b in zip(scene.objects, scene2.objects):
        assert float((a.pose() - b.pose()).norm()) < 1e-9                # tests/test_python.py:87
        assert float((a.mesh.pretransform - b.mesh.pretransform).norm()) < 1e-5
        assert (a.instance_index, a.mesh.class_index, a.casts_shadows) == (b.instance_index, b.mesh.class_index, b.casts_shadows)
        assert abs(a.metallic - b.metallic) < 1e-7 and abs(a.roughness - b.roughness) < 1e-7
    assert scene2.objects[0].mesh is scene2.objects[2].mesh and scene2.objects[0].mesh is not scene2.objects[1].mesh
    # a MeshCache makes deserialize() re-use meshes that are already loaded (and uploaded)
    cache = sl.MeshCache()
    cache.add(file_meshes)
    scene3 = sl.Scene((640, 480))
    scene3.deserialize(text, cache)
    assert all(o.mesh is file_meshes[k % 2] for k, o in enumerate(scene3.objects))
    with pytest.raises(RuntimeError, match="mesh subgroup"):
        sl.Scene((8, 8)).deserialize("[object]\npose=1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1\n")


@pytest.mark.gpu
def test_render_smoke_like_test_python():
    sl.init_cuda(0)
    scene = sl.Scene((640, 480))
    mesh = sl.Mesh.from_data(fixtures.load_mesh("bunny_mesh"))
    mesh.center_bbox()
    mesh.scale_to_bbox_diagonal(0.5)
    obj = sl.Object(mesh)
    scene.add_object(obj)
    pose = torch.eye(4)
    pose[2, 3] = 0.5
    obj.set_pose(pose)
    renderer = sl.RenderPass()
    result = renderer.render(scene)
    rgb = result.rgb()
    assert rgb.shape == (480, 640, 4) and rgb.dtype == torch.uint8 and rgb.is_cuda
    assert result.class_index().shape == (480, 640, 1) and result.class_index().dtype == torch.int16
    assert result.coordinates().shape == (480, 640, 3) and result.depth().shape == (480, 640)
    assert result.coordDepth().shape == (480, 640, 4) and result.normals().shape == (480, 640, 4)
    assert result.vertex_indices().shape == (480, 640, 3) and result.vertex_indices().dtype == torch.int32
    assert result.barycentric_coeffs().shape == (480, 640, 3) and result.cam_coordinates().shape == (480, 640, 4)
    assert (rgb[..., 3] != 0).sum() > 10
    # accessor tensors are fresh copies: a second render with result=None overwrites the pass's internal result
    depth_before = result.depth()
    pose[2, 3] = 0.8
    obj.set_pose(pose)
    result2 = renderer.render(scene)
    assert result2 is result
    assert not torch.equal(depth_before, result.depth())
    obj.instance_index = 0xFFFF
    assert int(renderer.render(scene).instance_index().min()) == -1      # int16 view of 65535


@pytest.mark.gpu
def test_sl_scene_matches_oracle():
    sl.init_cuda(0)
    cube = sl.Mesh.from_data(fixtures.load_mesh("cube_glb_mesh"))
    scene = sl.Scene((320, 240))
    obj = sl.Object(cube)
    scene.add_object(obj)
    scene.set_camera_look_at(torch.tensor([4.0, 0.0, 0.0]), torch.tensor([0.0, 0.0, 0.0]))
    scene.light_directions = torch.tensor([[-1.0, -1.0, -1.0], [0, 0, 0], [0, 0, 0]]) / np.sqrt(3.0)
    scene.light_colors = torch.tensor([[3.0, 3.0, 3.0], [0, 0, 0], [0, 0, 0]])
    scene.manual_exposure = 1.0
    scene.background_plane_size = torch.tensor([6.0, 6.0])
    plane_pose = torch.eye(4); plane_pose[2, 3] = -1.0
    scene.background_plane_pose = plane_pose
    rp = sl.RenderPass("flat")          # "flat" renders identically to "pbr" in the reference
    rp.ssao_enabled = False
    res = rp.render(scene, predicate=lambda o: True)
    got = {"rgb": res.rgb(), "coord": res.coordDepth(), "class_index": res.class_index(), "instance_index": res.instance_index(),
           "normals": res.normals(), "vertex_index": res._tensor(5), "barycentric": res._tensor(6), "cam_coord": res.cam_coordinates()}
    got = {k: v.cpu().numpy() for k, v in got.items()}
    got["class_index"] = got["class_index"].view(np.uint16)
    got["instance_index"] = got["instance_index"].view(np.uint16)
    got["vertex_index"] = got["vertex_index"].view(np.uint32)
    ref = ou.render(scene._spec(False, None), want_hdr=False)
    parity.assert_parity(got, ref, rgb_outliers=4)
    vi = got["vertex_index"][..., :3]
    assert len(np.unique(vi)) == 5                                   # background + 4 cube vertices (plane vertices carry id 0)
    hidden = rp.render(scene, predicate=lambda o: False)
    assert int(hidden.instance_index().max()) == 0
