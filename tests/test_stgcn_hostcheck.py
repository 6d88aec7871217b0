"""CPU check of the ST-GCN CUDA path's arithmetic: the per-element functions the kernels of
regennet_b200/csrc/stgcn.cu execute per thread, the packed-weight walk and the block schedule are compiled with g++
(tests/stgcn_hostcheck.cpp) and run on the host against the golden outputs of the imported reference and the oracle.
What this does NOT cover is the launch code itself (grids, streams, device allocations): see tests/test_gpu_stgcn.py."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import cases
from oracle import stgcn_ref
from regennet_b200 import _lib
from regennet_b200.stgcn import STGCN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ unavailable")
    out = str(tmp_path_factory.mktemp("stgcn") / "libstgcn_hostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "stgcn_hostcheck.cpp")])
    lib = ctypes.CDLL(out)
    lib.stgcn_host_packed_size.restype = ctypes.c_longlong
    lib.stgcn_host_packed_size.argtypes = [ctypes.POINTER(_lib.StgcnDesc)]
    lib.stgcn_host_forward.restype = ctypes.c_int
    lib.stgcn_host_forward.argtypes = [ctypes.POINTER(_lib.StgcnDesc), ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def _run(hostlib, model, c, x):
    desc = _lib.StgcnDesc(in_channels=c["in_channels"], num_person=c["num_person"], num_class=c["num_class"],
                          num_node=model.A.size(1), num_part=model.A.size(0))
    packed = model._packed(torch.device("cpu"))
    assert packed.numel() == hostlib.stgcn_host_packed_size(ctypes.byref(desc))
    N, T = x.shape[0], x.shape[3]
    feat = torch.empty(N, 256)
    yhat = torch.empty(N, c["num_class"])
    xc = x.contiguous()
    rc = hostlib.stgcn_host_forward(ctypes.byref(desc), packed.data_ptr(), packed.numel(), xc.data_ptr(), N, T,
                                    feat.data_ptr(), yhat.data_ptr())
    assert rc == 0, "host check returned %d" % rc
    return feat, yhat


@pytest.mark.parametrize("name", sorted(cases.STGCN_CASES))
def test_kernel_arithmetic_matches_reference_golden(hostlib, name):
    c = cases.STGCN_CASES[name]
    g = np.load(os.path.join(HERE, "stgcn.npz"))
    model = STGCN(in_channels=c["in_channels"], num_class=c["num_class"], num_person=c["num_person"],
                  graph_args=cases.stgcn_graph_args(c, ours=True), edge_importance_weighting=True, device="cpu")
    sd = stgcn_ref.make_state_dict(model.A.clone(), c["in_channels"], c["num_class"], c["num_person"], seed=c["wseed"])
    model.load_state_dict(sd, strict=True)
    x = torch.randn(c["N"], model.A.size(1), c["in_channels"], c["T"], generator=torch.Generator().manual_seed(c["xseed"]))
    feat, yhat = _run(hostlib, model, c, x)
    ef = np.abs(feat.numpy() - g[name + ".features"]).max()
    ey = np.abs(yhat.numpy() - g[name + ".yhat"]).max()
    print("%s: host-run kernel arithmetic vs reference golden: features %.3e yhat %.3e" % (name, ef, ey))
    assert ef < 1e-4 and ey < 1e-4


def test_kernel_arithmetic_chunking_and_odd_lengths(hostlib):
    """More samples than one chunk (64 (sample, person) rows) and lengths that the stride-2 blocks round up."""
    for P, N, T in [(2, 35, 9), (1, 70, 5), (2, 2, 1)]:
        c = dict(in_channels=6 * P, num_class=5, num_person=P)
        model = STGCN(in_channels=c["in_channels"], num_class=5, num_person=P,
                      graph_args={"layout": "openpose", "strategy": "spatial"}, edge_importance_weighting=True, device="cpu")
        sd = stgcn_ref.make_state_dict(model.A.clone(), c["in_channels"], 5, P, seed=3)
        model.load_state_dict(sd, strict=True)
        x = torch.randn(N, model.A.size(1), c["in_channels"], T, generator=torch.Generator().manual_seed(N))
        feat, yhat = _run(hostlib, model, c, x)
        with torch.no_grad():
            wf, wy = stgcn_ref.stgcn_forward(sd, x, P)
        assert (feat - wf).abs().max().item() < 1e-4 and (yhat - wy).abs().max().item() < 1e-4, (P, N, T)
