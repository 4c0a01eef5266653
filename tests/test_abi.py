"""The C-ABI library loads and exports every symbol include/ilf_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ilf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ilf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    for n in ("ilf_create", "ilf_destroy", "ilf_upload", "ilf_download", "ilf_set_deblock_info", "ilf_set_sao_params",
              "ilf_set_alf_params", "ilf_run", "ilf_deblock", "ilf_sao", "ilf_alf"):
        assert n in names


def test_library_exports_every_declared_symbol(ilf_lib):
    lib = ctypes.CDLL(ilf_lib.lib_path())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in ilf_b200.h but not exported: {missing}"
    assert lib.ilf_abi_version() == 1


def test_struct_sizes_match_header(ilf_lib):
    # ilf_sao_ctu is 32 bytes, ilf_alf_params 25*13*2 + 7*2 + 2, ilf_deblock_params 16 + 64*4
    import numpy as np
    import golden_io as G
    c = G.load_golden(G.golden_files()[0])
    assert c["sao_ctus"].shape[1] == 32
    assert c["alf_params"].size == 25 * 13 * 2 + 7 * 2 + 2
    assert c["db_params"].size == 16 + 64 * 4


def test_create_fails_loudly_without_gpu_or_with_bad_args(ilf_lib):
    import torch
    with pytest.raises(ilf_lib.IlfError):
        ilf_lib.InLoopFilter(417, 240)          # width not a multiple of 8
    if not torch.cuda.is_available():
        with pytest.raises(ilf_lib.IlfError, match="no CPU fallback"):
            ilf_lib.InLoopFilter(416, 240)
