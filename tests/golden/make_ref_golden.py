"""Generates golden vectors by RUNNING THE REFERENCE'S OWN CODE in the build container (it cannot travel to the GPU box,
the vectors do). Needs oracle/_ref (python oracle/build_ref.py) and /root/reference.

  golden_diff.npz   python/stillleben/diff.py (imported under a stub package: `Scene` / `RenderPassResult` fakes that serve
                    plain tensors) driving the reference's compiled extension oracle/_ref/diff (python/src/bridge_diff.cpp,
                    CPU branch — no GPU here): compute_image_space_gradients, dilate_object_mask per object,
                    backpropagate_gradient_to_poses, apply_pose_delta.
  cube_glb_mesh.npz / bunny_mesh.npz / kitchen_sink_mesh.npz / pbr_patch_mesh.npz
                    the reference's src/mesh_tools/consolidate.cpp + compute_tangents.cpp + vendored CgltfImporter run through
                    oracle/_ref/meshtool on the reference's two test assets and on tests/golden/assets/*.glb: the
                    consolidated 68-byte vertex stream, indices, sub-mesh table, materials as RenderShader::setMaterial
                    resolves them, textures with sampler state. tests/test_oracle_ref.py asserts stillleben_b200/gltf.py
                    reproduces them byte for byte.

    python tests/golden/make_ref_golden.py [diff] [mesh]
"""
import importlib.util
import os
import subprocess
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

REF = "/root/reference"
REFPY = os.path.join(REF, "python/stillleben")


class FakeObject:
    def __init__(self, pose, instance_index):
        self._pose, self.instance_index = torch.from_numpy(np.asarray(pose, np.float32)), int(instance_index)

    def pose(self):
        return self._pose


class FakeScene:
    def __init__(self, P, poses, ids):
        self._P = torch.from_numpy(np.asarray(P, np.float32))
        self.objects = [FakeObject(p, i) for p, i in zip(poses, ids)]

    def projection_matrix(self):
        return self._P


class FakeResult:
    """The accessor surface of RenderPassResult the diff module uses (py_render_pass.cpp:103-223)."""

    def __init__(self, rgb, inst, coord4):
        self._rgb, self._inst, self._c4 = torch.from_numpy(rgb), torch.from_numpy(inst), torch.from_numpy(coord4)

    def rgb(self):
        return self._rgb

    def instance_index(self):
        return self._inst.unsqueeze(-1)

    def coordinates(self):
        return self._c4[:, :, :3]

    def depth(self):
        return self._c4[:, :, 3]


def load_reference_diff():
    import build_ref
    ext = build_ref.load_diff()
    assert ext is not None, "run python oracle/build_ref.py diff first"
    pkg = types.ModuleType("slref"); pkg.__path__ = [REFPY]
    lib = types.ModuleType("slref.lib"); lib.__path__ = []
    core = types.ModuleType("slref.lib.libstillleben_python")
    core.Scene, core.RenderPassResult = FakeScene, FakeResult
    sys.modules.update({"slref": pkg, "slref.lib": lib, "slref.lib.libstillleben_python": core,
                        "slref.lib.libstillleben_diff_python": ext})
    for name in ("profiling", "diff"):
        spec = importlib.util.spec_from_file_location(f"slref.{name}", os.path.join(REFPY, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"slref.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["slref.diff"], ext


def diff_cases():
    import diff_ref
    import fixtures
    cases = [diff_ref.synthetic_inputs(seed, H=H, W=W) for seed, H, W in ((0, 60, 80), (1, 37, 131), (2, 64, 96))]
    # a real render: the committed oracle frame of the small table-top scene
    g = fixtures.load_golden("golden_tabletop")
    sc = fixtures.small_tabletop_scene()
    H, W = g["rgb"].shape[:2]
    grad = np.random.RandomState(5).normal(size=(3, H, W)).astype(np.float32)
    cases.append((g["rgb"], g["instance_index"].view(np.int16).reshape(H, W), g["coord"], grad, np.asarray(sc.projection, np.float32),
                  np.stack([np.asarray(o.pose, np.float32) for o in sc.objects]), np.array([o.instance_index for o in sc.objects], np.int32)))
    return cases


def make_diff():
    rd, ext = load_reference_diff()
    out = {}
    for i, (rgb, inst, coord4, grad, P, poses, ids) in enumerate(diff_cases()):
        scene, res = FakeScene(P, poses, ids), FakeResult(np.ascontiguousarray(rgb), np.ascontiguousarray(inst), np.ascontiguousarray(coord4))
        gx, gy, valid = rd.compute_image_space_gradients(scene, res)
        pg = rd.backpropagate_gradient_to_poses(scene, res, torch.from_numpy(grad))
        for k, v in (("rgb", rgb), ("inst", inst), ("coord4", coord4), ("grad", grad), ("P", P), ("poses", poses), ("ids", ids)):
            out[f"c{i}_{k}"] = v
        out[f"c{i}_grad_x"], out[f"c{i}_grad_y"] = gx.numpy(), gy.numpy()
        out[f"c{i}_valid"] = valid.numpy()
        out[f"c{i}_pose_grad"] = pg.numpy()
        dm, dc = [], []
        for idx in ids:
            m, c = ext.dilate_object_mask(torch.from_numpy(inst == idx), valid, torch.from_numpy(np.ascontiguousarray(coord4[:, :, :3])))
            dm.append(m.numpy()); dc.append(c.numpy())
        out[f"c{i}_dilated_mask"], out[f"c{i}_dilated_coords"] = np.stack(dm), np.stack(dc)
        print(f"diff case {i}: {inst.shape}, |pose_grad|max = {np.abs(pg.numpy()).max():.4g}")
    out["n_cases"] = np.array(len(diff_cases()))
    # apply_pose_delta (diff.py:525-590)
    rng = np.random.RandomState(11)
    poses = np.tile(np.eye(4, dtype=np.float32), (5, 1, 1))
    for b in range(5):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        poses[b, :3, :3] = q * np.sign(np.linalg.det(q))
        poses[b, :3, 3] = rng.normal(size=3)
    delta = (rng.normal(size=(5, 6)) * 0.05).astype(np.float32)
    out["apd_pose"], out["apd_delta"] = poses, delta
    out["apd_ortho"] = rd.apply_pose_delta(torch.from_numpy(poses), torch.from_numpy(delta)).numpy()
    out["apd_raw"] = rd.apply_pose_delta(torch.from_numpy(poses), torch.from_numpy(delta), orthonormalize=False).numpy()
    out["apd_single"] = rd.apply_pose_delta(torch.from_numpy(poses[2]), torch.from_numpy(delta[2])).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_diff.npz"), **out)
    print("wrote golden_diff.npz")


# Magnum sampler enums of the dump -> the ABI's (GL-style) enums
_WRAP = {0: 0, 1: 2, 2: 1, 3: 3}              # Repeat, MirroredRepeat, ClampToEdge, ClampToBorder -> abi.WRAP_*
_MINF = {(0, 0): 0, (1, 0): 1, (0, 1): 2, (1, 1): 3, (0, 2): 4, (1, 2): 5}     # (filter, mipmap) -> abi.FILTER_*


def ref_dump_to_mesh_npz(d, max_image=None):
    """The dump of oracle/_ref/meshtool in the fixture format of tests/fixtures.load_mesh (textures -> images with their
    sampler state, one per glTF texture, as Mesh::loadVisual creates one GL texture per TextureData: mesh.cpp:633-663)."""
    from PIL import Image
    pos = d["vertices"].view(np.float32).reshape(len(d["vertices"]), -1)[:, :3]
    out = {"vertices": d["vertices"], "indices": d["indices"], "submeshes": d["submeshes"],
           "bbox_min": pos.min(0).astype(np.float32), "bbox_max": pos.max(0).astype(np.float32),
           "materials": d["materials"][:, :15]}
    for t, (img, minf, mip, magf, ws, wt) in enumerate(d["textures"]):
        px = d[f"image{img}"]
        if max_image and px.shape[0] > max_image:        # keeps the fixture small (the bunny's 2048^2 texture)
            px = np.ascontiguousarray(np.asarray(Image.fromarray(px).resize((max_image, max_image), Image.BOX)))
        out[f"image{t}"] = px
        out[f"image{t}_sampler"] = np.array([_WRAP[int(ws)], _WRAP[int(wt)], _MINF[(int(minf), int(mip))], int(magf)])
    return out


def make_mesh():
    import hashlib
    import ref_meshdump
    tool = os.path.join(ROOT, "oracle", "_ref", "meshtool")
    assert os.path.exists(tool), "run python oracle/build_ref.py meshtool first"
    jobs = [(os.path.join(REF, "tests/cube.glb"), "cube_glb_mesh.npz", None),
            (os.path.join(REF, "tests/stanford_bunny/scene.gltf"), "bunny_mesh.npz", 256),
            (os.path.join(HERE, "assets/kitchen_sink.glb"), "kitchen_sink_mesh.npz", None),
            (os.path.join(HERE, "assets/pbr_patch.glb"), "pbr_patch_mesh.npz", None)]
    for src, dst, max_image in jobs:
        tmp = os.path.join("/tmp", dst + ".bin")
        subprocess.check_call([tool, src, tmp])
        d = ref_meshdump.read(tmp)
        out = ref_dump_to_mesh_npz(d, max_image)
        for k in [k for k in d if k.startswith("image")]:          # full-size image digests (the bunny texture is reduced above)
            out["sha256_" + k] = np.frombuffer(hashlib.sha256(d[k].tobytes()).digest(), np.uint8)
        np.savez_compressed(os.path.join(HERE, dst), **out)
        print("wrote", dst, {k: v.shape for k, v in out.items() if not k.startswith("sha")})


if __name__ == "__main__":
    want = sys.argv[1:] or ["diff", "mesh"]
    if "diff" in want:
        make_diff()
    if "mesh" in want:
        make_mesh()
