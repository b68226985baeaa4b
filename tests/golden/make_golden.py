"""Generates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only reads the .npz files):

  (cube_glb_mesh.npz / bunny_mesh.npz: the reference's two test assets after the reference's OWN consolidation — written by
   make_ref_golden.py, run it first)
  golden_cube.npz        oracle output for the reference's "vertex indices" test scene (tests/basic.cpp:375-453) at 320x240
  golden_bunny.npz       oracle output for the reference's "render" test scene (tests/basic.cpp:108-261) at 320x240, lit
  golden_tabletop.npz    oracle output of a small procedural table-top scene (regression pin)

The reference's tests hold no numeric golden vectors, so these vectors pin (a) the oracle
against the reference's own known-answer assertions on the reference's own assets and (b) the oracle
and the CUDA path against drift.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_util as ou  # noqa: E402
import fixtures  # noqa: E402
from stillleben_b200 import gltf  # noqa: E402

REF = "/root/reference/tests"


def save_mesh(path, m):
    d = {"vertices": m.vertices.view(np.uint8).reshape(len(m.vertices), -1), "indices": m.indices,
         "submeshes": np.array(m.submeshes, np.int64), "bbox_min": m.bbox_min, "bbox_max": m.bbox_max,
         "materials": np.array([[*x.base_color, *x.emissive, x.metallic, x.roughness, x.tex_base_color, x.tex_normal,
                                 x.tex_metallic_roughness, x.tex_emissive, x.tex_occlusion] for x in m.materials], np.float64)}
    for i, im in enumerate(m.images):
        d[f"image{i}"] = im.pixels
        d[f"image{i}_sampler"] = np.array([im.wrap_s, im.wrap_t, im.min_filter, im.mag_filter])
    np.savez_compressed(path, **d)


def save_frame(path, frame):
    np.savez_compressed(path, **frame)


def main():
    # cube_glb_mesh.npz / bunny_mesh.npz are written by make_ref_golden.py from the REFERENCE's own consolidation
    assets = ou.OracleAssets()
    for name, scene in (("golden_cube", fixtures.cube_test_scene(320, 240)), ("golden_bunny", fixtures.bunny_test_scene(320, 240, lit=True)),
                        ("golden_tabletop", fixtures.small_tabletop_scene())):
        out = ou.render(scene, assets)
        save_frame(os.path.join(HERE, name + ".npz"), out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
