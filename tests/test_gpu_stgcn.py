"""GPU parity of the evaluation feature extractor (ST-GCN) through the public mirror
``regennet_b200.stgcn.STGCN`` -> regen_stgcn_* (regennet_b200/csrc/stgcn.cu) vs the golden outputs of the imported
reference (tests/golden/make_golden_stgcn.py) and the oracle.

The per-element arithmetic, the packed-weight walk, the block schedule, chunking and workspace sizing are additionally
verified on the CPU by tests/test_stgcn_hostcheck.py."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import stgcn_ref
from regennet_b200.stgcn import STGCN

pytestmark = pytest.mark.gpu
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4   # fp32 CUDA-core arithmetic, summation order differs from torch's convolutions


def _model(c, layout, seed):
    m = STGCN(in_channels=c["in_channels"], num_class=c["num_class"], num_person=c["num_person"],
              graph_args=cases.stgcn_graph_args(dict(c, layout=layout), ours=True), edge_importance_weighting=True,
              device="cuda")
    sd = stgcn_ref.make_state_dict(m.A.clone(), c["in_channels"], c["num_class"], c["num_person"], seed=seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("name", sorted(cases.STGCN_CASES))
def test_stgcn_matches_reference_golden(built_lib, name):
    c = cases.STGCN_CASES[name]
    g = np.load(os.path.join(HERE, "stgcn.npz"))
    model, sd = _model(c, c["layout"], c["wseed"])
    x = torch.randn(c["N"], model.A.size(1), c["in_channels"], c["T"], generator=torch.Generator().manual_seed(c["xseed"]))
    launches0 = built_lib.regen_launch_count()
    with torch.no_grad():
        batch = model({"output": x.cuda()})
    assert built_lib.regen_launch_count() > launches0
    ef = np.abs(batch["features"].cpu().numpy() - g[name + ".features"]).max()
    ey = np.abs(batch["yhat"].cpu().numpy() - g[name + ".yhat"]).max()
    print("%s: max abs err vs reference golden: features %.3e yhat %.3e" % (name, ef, ey))
    assert ef < TOL and ey < TOL


def test_stgcn_chunking_and_odd_lengths_match_oracle(built_lib):
    for P, N, T in [(2, 35, 9), (1, 70, 5), (2, 2, 1), (2, 100, 60)]:
        c = dict(in_channels=6 * P, num_class=5, num_person=P)
        model, sd = _model(c, "openpose", 3)
        x = torch.randn(N, model.A.size(1), c["in_channels"], T, generator=torch.Generator().manual_seed(N))
        with torch.no_grad():
            batch = model({"output": x.cuda()})
            sel = slice(0, N) if N <= 70 else slice(60, 70)      # rows of the second chunk at the large size
            wf, wy = stgcn_ref.stgcn_forward(sd, x[sel], P)
        assert (batch["features"].cpu()[sel] - wf).abs().max().item() < TOL, (P, N, T)
        assert (batch["yhat"].cpu()[sel] - wy).abs().max().item() < TOL, (P, N, T)


def test_stgcn_full_eval_shape_matches_oracle(built_lib):
    """The evaluation's real shape -- SMPL-X-shaped graph (56 nodes), two persons, T = 60, more samples than one chunk --
    against the oracle on the rows around the chunk boundary."""
    c = dict(cases.STGCN_CASES["stgcn_smplx_p2"])
    model, sd = _model(c, c["layout"], 9)
    N = 70
    x = torch.randn(N, 56, 12, 60, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        batch = model({"output": x.cuda()})
        sel = slice(28, 36)                                      # chunks hold 32 samples of two persons
        wf, wy = stgcn_ref.stgcn_forward(sd, x[sel], 2)
    assert (batch["features"].cpu()[sel] - wf).abs().max().item() < TOL
    assert (batch["yhat"].cpu()[sel] - wy).abs().max().item() < TOL


def test_stgcn_error_behaviour(built_lib):
    c = dict(in_channels=12, num_class=5, num_person=2)
    model, _ = _model(c, "openpose", 0)
    with pytest.raises(ValueError):
        model({"output": torch.zeros(2, 7, 12, 8, device="cuda")})            # wrong joint count
    with pytest.raises(RuntimeError):
        model({"output": torch.zeros(2, 18, 12, 8)})                          # CPU tensor: no fallback
    model.train()
    with pytest.raises(NotImplementedError):
        model({"output": torch.zeros(2, 18, 12, 8, device="cuda")})
