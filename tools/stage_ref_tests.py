"""Stage the reference's OWN Python tests / example next to the test-suite so they can run, unmodified, on the GPU box.

/root/reference does not exist on the GPU box and reference sources must not enter the repository, so — like oracle/_ref —
the files are copied by this recipe into tests/_ref/ (git-ignored, NOT gpurun-ignored: it travels with the snapshot):
    tests/test_python.py, tests/test_grad.py, tests/stanford_bunny/, tests/cube.glb, examples/ycb.py,
    pbr_root/examples/pbr.py + pbr_root/tests/stanford_bunny/ (pbr.py finds its assets relative to its own location)
tests/test_gpu_reference_suite.py runs them against the `stillleben` package of this repository; it skips when the
directory was never staged.        python tools/stage_ref_tests.py [/root/reference]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage(ref="/root/reference"):
    dst = os.path.join(ROOT, "tests", "_ref")
    if not os.path.isdir(os.path.join(ref, "tests")):
        return False
    os.makedirs(dst, exist_ok=True)
    for rel in ("tests/test_python.py", "tests/test_grad.py", "tests/cube.glb", "examples/ycb.py"):
        shutil.copyfile(os.path.join(ref, rel), os.path.join(dst, os.path.basename(rel)))
    bunny = os.path.join(dst, "stanford_bunny")
    shutil.rmtree(bunny, ignore_errors=True)
    shutil.copytree(os.path.join(ref, "tests", "stanford_bunny"), bunny)
    pbr = os.path.join(dst, "pbr_root")
    shutil.rmtree(pbr, ignore_errors=True)
    os.makedirs(os.path.join(pbr, "examples"))
    shutil.copyfile(os.path.join(ref, "examples", "pbr.py"), os.path.join(pbr, "examples", "pbr.py"))
    shutil.copytree(os.path.join(ref, "tests", "stanford_bunny"), os.path.join(pbr, "tests", "stanford_bunny"))
    return True


if __name__ == "__main__":
    ok = stage(*sys.argv[1:2])
    print("staged tests/_ref" if ok else "reference not found; nothing staged")
