"""CPU tests of the boundary: the C-ABI library builds, loads, exports every symbol include/mmrecall.h declares,
and fails loudly (no fallback) without an sm_100 device.  No compute calls here."""
import ctypes
import os
import re

import pytest

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mmrecall.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.EXPORTS)


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load(build_if_missing=True)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"libmmrecall.so does not export {name}"
    assert lib.mmr_abi_version() == 2


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.MmrConfig) == 17 * 4
    assert ctypes.sizeof(_lib.MmrInputs) == 14 * ctypes.sizeof(ctypes.c_void_p)   # 13 pointers + int32 (padded)
    assert ctypes.sizeof(_lib.MmrTensor) == 8 + 8 + 8 + 4 * 8   # name, data, ndim (+pad), dims[4]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only machines")
    lib = _lib.load(build_if_missing=True)
    assert lib.mmr_device_check(0) == 3   # MMR_ERR_ARCH
    assert b"no CPU fallback" in lib.mmr_last_error()
    with pytest.raises(_lib.MmrError):
        _lib.check(lib.mmr_cast16(0, 0, 8, _lib.DT_FP16, 0))
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, ModelConfig
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    cfg = ModelConfig(LDS, n_layers=1, lq=4, nbox=2, vocab=50)
    with pytest.raises(_lib.MmrError):
        MatchScorer(cfg, synth.make_weights(cfg, seed=1), device=0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kddcup_2020_multimodalitiesrecall_2nd_place_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_workspace_bytes_is_a_host_side_plan():
    """mmr_workspace_bytes needs no GPU: cfg2 (12-layer zk, 32 x 36, batch 256) plans ~0.75 GB of weights (0.52 GB of it
    the label-conv tables) and ~0.6 GB of workspace; strict precision triples the matrices and swaps the activations."""
    lib = _lib.load(build_if_missing=True)

    def plan(kind, precision, batch=256, n_layers=12, r=0, x=0):
        c = _lib.MmrConfig(model_kind=kind, dtype=0, hidden=768, heads=12, intermediate=3072, vocab=21128, max_pos=512,
                           type_vocab=2, feat_dim=2048, label_len=8, n_layers=n_layers, n_r_layers=r, n_x_layers=x, lq=32,
                           nbox=36, max_batch=batch, precision=precision)
        w, ws = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(lib.mmr_workspace_bytes(ctypes.byref(c), ctypes.byref(w), ctypes.byref(ws)))
        return w.value, ws.value
    w, ws = plan(_lib.MODEL_ZK, 0)
    assert 0.7e9 < w < 0.9e9 and 0.4e9 < ws < 0.9e9
    w2, ws2 = plan(_lib.MODEL_ZK, 1)
    assert w2 > w + 0.3e9 and ws2 > ws
    assert plan(_lib.MODEL_ZK, 0, batch=128)[1] < ws
    wl, _ = plan(_lib.MODEL_LXMERT, 0, n_layers=9, r=5, x=5)
    assert 0.3e9 < wl < 0.6e9
