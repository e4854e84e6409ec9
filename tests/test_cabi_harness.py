"""The C ABI driven from COMPILED C (gcc, links libzgpu.so): the INTEGRATION.md snippet, 32 concurrent
zg_check_bulk_str callers with per-thread error text, the ZG_E2BIG sizing protocol. Every other test goes
through ctypes; this one is what a cgo caller sees (SURVEY.md 8b "Errors" / "Threading")."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "spicedb-kubeapi-proxy_b200")


def _build(tmp_path):
    import zgpu

    zgpu.build_library()
    exe = str(tmp_path / "cabi_harness")
    subprocess.run(["gcc", "-O1", "-Wall", "-Wextra", "-std=c11", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cabi", "harness.c"), "-o", exe, "-L", PKG, "-lzgpu",
                    "-Wl,-rpath," + PKG, "-lpthread"], check=True)
    return exe


def test_harness_compiles_and_links_against_the_header(tmp_path):
    """CPU: the header is valid C11 and the library exports what the harness uses."""
    _build(tmp_path)


@pytest.mark.gpu
def test_compiled_c_caller_concurrent_and_exact(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cabi harness ok" in r.stdout
