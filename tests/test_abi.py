"""The C-ABI library loads on a machine without a GPU and exports every symbol include/slb.h declares;
the ctypes mirror has the same struct layouts as the C header. No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from stillleben_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slb.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(lib.LIB_PATH), "build with __graft_entry__.build() first"
    L = C.CDLL(lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in slb.h but not exported"


def test_ctypes_mirror_covers_header():
    assert sorted(abi.PROTOTYPES) == declared_symbols()
    abi.bind(C.CDLL(lib.LIB_PATH))
    assert lib.load().slb_abi_version() == abi.SLB_ABI_VERSION


def test_struct_layouts_match_c():
    prog = r'''
#include <stdio.h>
#include "slb.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(slb_image), sizeof(slb_submesh), sizeof(slb_material), sizeof(slb_lightmap_desc),
         sizeof(slb_object_desc), sizeof(slb_scene_desc), sizeof(slb_stats));
  printf("%zu %zu %zu\n", __builtin_offsetof(slb_scene_desc, objects), __builtin_offsetof(slb_object_desc, sticker_range),
         __builtin_offsetof(slb_scene_desc, manual_exposure));
  printf("%zu %zu %zu\n", sizeof(slb_camera_params), __builtin_offsetof(slb_camera_params, stages), __builtin_offsetof(slb_camera_params, seed));
  return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    sizes = [int(x) for x in out]
    mirror = [C.sizeof(abi.Image), C.sizeof(abi.Submesh), C.sizeof(abi.Material), C.sizeof(abi.LightmapDesc),
              C.sizeof(abi.ObjectDesc), C.sizeof(abi.SceneDesc), C.sizeof(abi.Stats),
              abi.SceneDesc.objects.offset, abi.ObjectDesc.sticker_range.offset, abi.SceneDesc.manual_exposure.offset,
              C.sizeof(abi.CameraParams), abi.CameraParams.stages.offset, abi.CameraParams.seed.offset]
    assert sizes == mirror


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.SlbError):
        lib.Context(0)
