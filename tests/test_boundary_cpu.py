"""CPU-side checks of the drop-in boundary: module surface, state-dict layout, the stock (autograd)
route against the reference's golden outputs, and that the C-ABI library loads and exports every
symbol include/cdl_b200.h declares.  No GPU compute here."""
import copy
import ctypes
import os
import pickle
import re

import numpy as np
import pytest
import torch

import cdlnet_video_b200 as cb
from cdlnet_video_b200 import _lib
from util import CASES, load_case, module_from_case, case_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session", autouse=True)
def built():
    _lib.build()


def test_header_symbols_exported():
    """Every function declared in include/cdl_b200.h is exported by libcdl_b200.so (and vice versa for the loader)."""
    hdr = open(os.path.join(ROOT, "include", "cdl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cdl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS)


def test_library_loads_and_reports_errors_without_gpu():
    lib = _lib.load()
    assert lib.cdl_abi_version() == 1
    assert lib.cdl_status_string(0) == b"ok"
    assert b"not supported" in lib.cdl_status_string(-3)
    handle = ctypes.c_void_p()
    assert lib.cdl_plan_create(ctypes.byref(handle), None) == -1           # CDL_ERR_NULL
    d = _lib.CdlDesc()
    d.ndim = 4
    assert lib.cdl_plan_create(ctypes.byref(handle), ctypes.byref(d)) == -2   # CDL_ERR_SHAPE
    d.ndim, d.N, d.C, d.M, d.K, d.s = 2, 1, 1, 8, 2, 2
    d.dims[:] = [1, 16, 16]
    d.P[:] = [1, 6, 6]
    assert lib.cdl_plan_create(ctypes.byref(handle), ctypes.byref(d)) == -3   # even filter: unsupported
    if not torch.cuda.is_available():
        d.P[:] = [1, 7, 7]
        assert lib.cdl_plan_create(ctypes.byref(handle), ctypes.byref(d)) == -6   # no device: loud, no CPU fallback
        with pytest.raises(RuntimeError, match="no usable CUDA device"):
            cb.Plan(2, 1, 1, 8, 2, (16, 16), (7, 7), 2)


def test_desc_struct_matches_header():
    hdr = open(os.path.join(ROOT, "include", "cdl_b200.h")).read()
    body = re.search(r"typedef struct cdl_desc \{(.*?)\} cdl_desc_t;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    n_ints = 0
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        assert decl.startswith("int32_t")
        for item in decl[len("int32_t"):].split(","):
            m = re.search(r"\[(\d+)\]", item)
            n_ints += int(m.group(1)) if m else 1
    assert ctypes.sizeof(_lib.CdlDesc) == 4 * n_ints


@pytest.mark.parametrize("name", CASES)
def test_stock_route_matches_reference_golden(name):
    """The autograd/CPU route of the drop-in modules reproduces the reference bit-for-bit-ish (same torch ops)."""
    d = load_case(name)
    net = module_from_case(d, name)
    y, sigma, mask = case_inputs(d)
    with torch.no_grad():
        xhat, z = net(y, sigma, mask=mask)
    assert xhat.shape == d["xhat"].shape and z.shape == d["z"].shape
    assert np.abs(xhat.numpy() - d["xhat"]).max() <= 1e-6
    assert np.abs(z.numpy() - d["z"]).max() <= 1e-6
    if "trace" in d:
        with torch.no_grad():
            items = list(net.forward_generator(y, sigma, mask=mask))
        assert len(items) == net.K + 1
        for k in range(net.K):
            assert np.abs(items[k].numpy() - d["trace"][k]).max() <= 1e-6
        assert np.abs(items[-1].numpy() - d["xhat"]).max() <= 1e-6


def test_state_dict_keys_match_reference_layout():
    """2D `t, g, A.k.weight, B.k.weight, D.weight`; 3D without g; GDLNet 29 keys at K=3."""
    n2 = cb.CDLNet(K=3, M=4, P=7, s=2, init=False)
    assert set(n2.state_dict()) == {"t", "g", "D.weight"} | {f"{s}.{k}.weight" for s in "AB" for k in range(3)}
    assert n2.t.shape == (3, 2, 4, 1, 1) and n2.D is n2.B[0]
    n3 = cb.CDLNetVideo(K=2, M=4, P=7, s=2, init=False)
    assert set(n3.state_dict()) == {"t", "D.weight"} | {f"{s}.{k}.weight" for s in "AB" for k in range(2)}
    assert n3.A[0].weight.shape == (4, 1, 7, 7, 7) and n3.t.shape == (2, 2, 4, 1, 1, 1)     # int P -> cubic (F4)
    n3b = cb.CDLNetVideo(K=1, M=4, P=[9, 9, 5], s=2, init=False)
    assert n3b.A[0].weight.shape == (4, 1, 9, 9, 5) and n3b.A[0].padding == (4, 4, 2)        # (frames, rows, cols) (F5)
    g = cb.GDLNet(K=3, M=4, P=7, s=2, C=3, order=2, init=False)
    assert len(g.state_dict()) == 29
    assert g.A[0].alpha.shape == (2, 4, 3, 1, 1)
    # weights are shared A <-> B at construction (adjoint pair)
    assert torch.equal(n2.A[1].weight, n2.B[2].weight)


def test_module_survives_pickle_and_deepcopy():
    net = cb.CDLNet(K=2, M=4, P=7, s=1, init=False)
    net.__dict__["_plans"] = {"x": ctypes.c_void_p(1)}     # stand-in for live native handles
    net2 = pickle.loads(pickle.dumps(net))
    assert "_plans" not in net2.__dict__
    net3 = copy.deepcopy(net)
    assert torch.equal(net3.t, net.t)


def test_project_and_init_power_method():
    torch.manual_seed(0)
    net = cb.CDLNet(K=2, M=8, P=7, s=2, C=1, t0=-0.1, init=True)       # runs the power method (CPU, stock convs)
    # after spectral normalisation the largest eigenvalue of D∘A is ~1
    x = torch.rand(1, 1, 64, 64)
    with torch.no_grad():
        for _ in range(50):
            x = net.D(net.A[0](x))
            x = x / x.norm()
        L = float((x * net.D(net.A[0](x))).sum())
    assert 0.9 < L < 1.1
    net.project()
    assert float(net.t.min()) == 0.0
    assert float(torch.norm(net.A[0].weight, dim=(2, 3)).max()) <= 1.0 + 1e-6


def test_grad_mode_uses_differentiable_route():
    net = cb.CDLNet(K=2, M=4, P=7, s=2, t0=0.01, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(2):
            net.A[k].weight.mul_(0.05)
            net.B[k].weight.mul_(0.05)
    y = torch.rand(1, 1, 16, 16)
    xhat, z = net(y, 25.0)
    xhat.sum().backward()
    assert net.t.grad is not None and net.A[1].weight.grad is not None


def test_drop_in_import_as_model_net():
    """With the package directory first on sys.path the classes are importable as `model.net.*`
    (what the reference's train/analyze drivers import)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import model.net as mn; "
            "n = mn.CDLNetVideo(K=1, M=2, P=7, s=2, init=False); print(type(n).__module__, n.K)"
            % os.path.join(ROOT, "cdlnet-video_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/")
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["model.net", "1"]
