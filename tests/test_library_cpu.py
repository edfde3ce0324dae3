"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every declared symbol;
the host-side guards fail loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from npvp_b200.build import build
    return build()


def test_header_symbols_exported(lib_path):
    from npvp_b200 import _lib
    header = open(os.path.join(ROOT, "include", "npvp_b200.h")).read()
    declared = set(re.findall(r"\b(npvp_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(lib_path)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/npvp_b200.h but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib.AUX_SYMBOLS)
    assert declared == bound, f"header vs ctypes table mismatch: {declared ^ bound}"
    assert _lib.load_library(lib_path).npvp_version() == 100


def test_argument_validation_without_gpu(lib_path):
    from npvp_b200 import _lib
    lib = _lib.load_library(lib_path)
    ep = _lib.Epilogue()
    rc = lib.npvp_gemm_bf16(None, 0, None, 0, 0, 0, 0, ctypes.byref(ep), 0, None)
    assert rc == -1 and b"null pointer" in lib.npvp_last_error()


def test_no_cpu_fallback():
    import npvp_b200
    hl = torch.linspace(0, 7, 8)
    pred = npvp_b200.Predictor(8, 8, 4, hl, hl, torch.tensor([0., 1.]), torch.tensor([2., 3.]), 512, 'Add', 'layer', 256, 1,
                               False, 1, evt_former_num_layers=1).eval()
    with pytest.raises(NotImplementedError, match="CUDA"):
        pred(torch.zeros(1, 2, 512, 8, 8))
    pred.train()
    with pytest.raises(NotImplementedError, match="inference"):
        pred(torch.zeros(1, 2, 512, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA device"):
        _ops = __import__("npvp_b200._lib", fromlist=["Ops"]).Ops()
        _ops.layernorm_rows(torch.zeros(4, 512), torch.ones(512), torch.zeros(512), out_f32=torch.zeros(4, 512))
    with pytest.raises(NotImplementedError):
        npvp_b200.ResnetEncoder(3, learn_3d=True)
    with pytest.raises(ValueError):
        npvp_b200.ResnetDecoder(3, out_layer="Softmax")
    with pytest.raises(AssertionError, match="Invalid T"):
        pred.reset_pos_coor(torch.tensor([0., 1.]), torch.tensor([9.]))


def test_config_presets_and_timestamps():
    from npvp_b200.config import preset
    from npvp_b200.pipeline import timestamp_lists
    cfg = preset("KITTI_VFP_NPVP-S")
    to, tp = timestamp_lists(cfg)
    assert to.tolist() == [0, 1, 2, 3] and tp.tolist() == [4, 5, 6, 7, 8]
    ref_yaml = "/root/reference/configs/config_KITTI_VFP_NPVP-S.yaml"
    if os.path.exists(ref_yaml):          # only in the build container
        from npvp_b200.config import load_config
        y = load_config(ref_yaml)
        assert y.Predictor.max_T == cfg.Predictor.max_T and y.AE.ngf == cfg.AE.ngf and isinstance(y.AE.AE_lr, float)
