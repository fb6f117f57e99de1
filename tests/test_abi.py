"""CPU-side checks of the drop-in boundary: libb200rl.so loads without a GPU, exports every
symbol include/b200rl.h declares, the ctypes table covers the header, and the product path
fails loudly (no CPU fallback) when no device is present."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200rl.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(pkg):
    from b200rl import _lib
    if not os.path.exists(_lib.SO_PATH):
        import __graft_entry__ as g
        g.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH], text=True)
    exported = set(re.findall(r" T (b200rl_[a-z0-9_]+)", out))
    declared = _declared()
    assert len(declared) > 30
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in b200rl.h but not exported: {missing}"


def test_ctypes_table_matches_header(pkg):
    from b200rl import _lib
    lib = _lib.load()  # dlopen works without a GPU
    declared = _declared()
    assert sorted(_lib.SIGNATURES) == declared
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b200rl_abi_version() >= 1


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import/load it."""
    pkg_dir = os.path.join(ROOT, "reinforcementlearning.jl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "oracle/" not in txt, f


def test_no_device_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B200RLError) as ei:
        pkg.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_pure_c_client_builds_links_and_fails_loudly_without_a_gpu(pkg, tmp_path):
    """examples/ppo_cartpole.c is what any FFI does: C99, include/b200rl.h only.  It must compile and link against the
    in-tree library; without a device the first call reports the error through the ABI's convention (no fallback)."""
    import subprocess
    exe = str(tmp_path / "ppo_cartpole")
    libdir = os.path.dirname(pkg._lib.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "ppo_cartpole.c"), "-L", libdir, "-lb200rl", "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = subprocess.run([exe, "256", "1"], capture_output=True, text=True, timeout=120)
    if has_gpu:
        assert r.returncode == 0 and "env-steps/s" in r.stdout, r.stderr
    else:
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
