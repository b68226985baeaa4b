"""The compiled C++ boundary: include/stillleben_shim.hpp gives a C++ caller the reference's sl::Context / Mesh / Object / Scene /
RenderPass / RenderPass::Result names over the C ABI; tests/cpp/shim_client.cpp replays the reference's "vertex indices" and
"render" test cases (tests/basic.cpp:108-261,375-453) through it. Built by tests/cpp/Makefile (g++ only, links libslb.so)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_client")


def test_shim_header_compiles_and_links():
    """No GPU needed: the header is valid C++17 against include/slb.h and every ABI symbol it uses resolves at link time."""
    subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "tests", "cpp")])
    assert os.access(EXE, os.X_OK)


@pytest.mark.gpu
def test_cpp_client_replays_the_reference_test_cases():
    if not os.access(EXE, os.X_OK):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL CHECKS PASSED" in r.stdout, r.stdout + r.stderr
