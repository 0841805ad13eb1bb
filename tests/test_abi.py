"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/madrona_b200.h declares; the product path fails loudly without CUDA."""
import os
import re

import pytest

import madrona_b200 as mb
from madrona_b200.executor import EXPORTED_SYMBOLS


def _header_symbols(root):
    text = open(os.path.join(root, "include", "madrona_b200.h")).read()
    return sorted(set(re.findall(r"\b(mb2_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(repo_root):
    lib = mb.load_library()
    declared = _header_symbols(repo_root)
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert sorted(EXPORTED_SYMBOLS) == declared


def test_version_string():
    assert b"sm_100a" in mb.load_library().mb2_version()


def test_executor_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sims import make_executor
    with pytest.raises(mb.MadronaB200Error):
        make_executor("cartpole", 4)


def test_jit_precompile_reports_errors(tmp_path):
    bad = tmp_path / "bad.cpp"
    bad.write_text("#include <madrona/custom_context.hpp>\nthis is not c++\n")
    with pytest.raises(mb.MadronaB200Error) as e:
        mb.precompile(mb.CompileConfig([str(bad)]))
    assert "NVRTC" in str(e.value)
