"""``model.rot2xyz(...)`` drop-in: regennet_b200.rotation2xyz.Rotation2xyz vs goldens written by the
reference's Rotation2xyz / Rotation2xyz_x (tests/golden/make_golden_rot2xyz.py; smplx layers replaced by a stub with the
same call contract on both sides).

CPU tests check the host logic (person split, mask scatter, re-rooting, translation, pose slicing) with the rot6d kernel
replaced by the oracle's restatement; GPU tests run the product path (library kernel) and the call sites' exact keyword
sets (sample/cgenerate.py:156-158, eval/a2m/stgcn_eval.py:81-83)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import sampler_ref
from regennet_b200 import rotation2xyz as r2x_mod
from regennet_b200.rotation2xyz import Rotation2xyz

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = np.load(os.path.join(HERE, "rot2xyz.npz"))


def _call(r2x, c, x, mask, back=False):
    return r2x(x=x, mask=mask, pose_rep="rot6d", glob=c["glob"], translation=True, jointstype=c["jointstype"],
               vertstrans=True, num_person=c["P"], betas=None, beta=0,
               glob_rot=None if c["glob"] else cases.ROT2XYZ_GLOB_ROT, get_rotations_back=back)


def _make(c, device):
    r = Rotation2xyz(device="cpu", dataset="ntu", body_model=c["body_model"])
    assert r.smpl_model is None          # no smplx package / files in this image
    r.attach_body_model(cases.stub_body_model(c["body_model"]).to(device))
    return r


@pytest.mark.parametrize("name", sorted(cases.ROT2XYZ_CASES))
def test_host_logic_matches_reference_golden(name, monkeypatch):
    c = cases.ROT2XYZ_CASES[name]
    monkeypatch.setattr(r2x_mod.geometry, "rotation_6d_to_matrix", sampler_ref.rotation_6d_to_matrix)
    x, mask = cases.rot2xyz_inputs(c)
    got = _call(_make(c, "cpu"), c, x, mask)
    assert tuple(got.shape) == GOLD[name].shape
    assert np.abs(got.numpy() - GOLD[name]).max() < 1e-5
    if name + "/rotations" in GOLD.files:
        _, rot, go = _call(_make(c, "cpu"), c, x, mask, back=True)
        assert np.abs(rot.numpy() - GOLD[name + "/rotations"]).max() < 1e-6
        assert np.abs(go.numpy() - GOLD[name + "/global_orient"]).max() < 1e-6


def test_error_behaviour_matches_reference():
    c = cases.ROT2XYZ_CASES["x_p1_mask"]
    x, mask = cases.rot2xyz_inputs(c)
    r = Rotation2xyz(body_model="smplx")
    assert r(x, mask, "xyz", True, True, "smplx", True) is x                      # model/rotation2xyz.py:172-173
    with pytest.raises(TypeError):                                               # :178-179
        r(x, mask, "rot6d", True, False, "smplx", True, glob_rot=None)
    with pytest.raises(NotImplementedError):                                     # :181-182
        r(x, mask, "rot6d", True, True, "nope", True)
    with pytest.raises(RuntimeError, match="body model"):                        # nothing attached
        r(x, mask, "rot6d", True, True, "smplx", True)


def test_cmdm_moves_attached_body_model():
    """model/cmdm.py:255-262: .to() / .train() reach the body model through rot2xyz."""
    from regennet_b200.cmdm import CMDM
    m = CMDM(**dict(cases.MODELS["ntu"], num_layers=1))
    body = cases.stub_body_model("smplx")
    m.rot2xyz.attach_body_model(body)
    m.double()
    assert body.rest.dtype == torch.float64
    m.float().train()
    assert body.training
    m.eval()
    assert not body.training and body.rest.dtype == torch.float32


# --------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.ROT2XYZ_CASES))
def test_gpu_call_matches_reference_golden(built_lib, name):
    c = cases.ROT2XYZ_CASES[name]
    x, mask = cases.rot2xyz_inputs(c)
    n0 = built_lib.regen_launch_count()
    got = _call(_make(c, "cuda"), c, x.cuda(), mask.cuda())
    assert built_lib.regen_launch_count() > n0            # the rot6d kernel ran
    err = np.abs(got.cpu().numpy() - GOLD[name]).max()
    print("%s: max abs err vs reference golden %.3e" % (name, err))
    assert err < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("P", [1, 2])
def test_gpu_rotations_vs_oracle(built_lib, P):
    """.rotations(): per person, masked frames only, translation row dropped -- vs the oracle's rot6d -> rotmat."""
    c = dict(B=4, J=56, P=P, T=11, cut=[0, 5, 10, 1], seed=70 + P)
    x, mask = cases.rot2xyz_inputs(c)
    r = Rotation2xyz(body_model="smplx")
    out = r.rotations(x.cuda(), mask.cuda(), pose_rep="rot6d", translation=True, num_person=P)
    assert len(out) == P
    for pid, o in enumerate(out):
        xp = x[:, :, 6 * pid:6 * pid + 6]
        want = sampler_ref.rotation_6d_to_matrix(xp[:, :-1].permute(0, 3, 1, 2)[mask])
        assert tuple(o["rotations"].shape) == (int(mask.sum()), 55, 3, 3)
        assert (o["rotations"].cpu() - want).abs().max() < 1e-5
        assert torch.equal(o["translations"].cpu(), xp[:, -1, :3])
    # mask=None selects every frame
    full = r.rotations(x.cuda(), None, num_person=P)
    assert full[0]["rotations"].shape[0] == c["B"] * c["T"]


@pytest.mark.gpu
def test_gpu_cgenerate_loop_body(built_lib):
    """The body of the reference's sampling script (sample/cgenerate.py:121-158) against this package's classes with the
    script's exact keyword sets: p_sample_loop(progress=True, dump_steps=None, noise=None, const_noise=False, ...),
    scipy gaussian_filter1d round trip, then model.rot2xyz(...) with a body model attached."""
    from scipy.ndimage import gaussian_filter1d
    from regennet_b200 import synthetic
    from regennet_b200.cfg_sampler import ClassifierFreeSampleModel
    from regennet_b200.cmdm import CMDM
    from test_gpu_sampler import _diffusion
    mk = dict(cases.MODELS["chi3d"], num_layers=2)
    model = CMDM(**mk)
    model.load_state_dict(synthetic.make_state_dict(seed=3, **dict(cases.synth_kw("chi3d"), num_layers=2)), strict=False)
    model.rot2xyz.attach_body_model(cases.stub_body_model("smplx"))
    model = ClassifierFreeSampleModel(model.cuda().eval())
    diffusion = _diffusion("ddim4")
    batch_size, n_frames = 3, 20
    _, y = synthetic.make_inputs(batch_size, 56, 6, n_frames, seed=5, cond_mode="action", num_actions=8)
    model_kwargs = {"y": {k: v.cuda() for k, v in y.items()}}
    lengths = torch.tensor([20, 13, 7])
    model_kwargs["y"]["lengths"] = lengths.cuda()
    model_kwargs["y"]["mask"] = (torch.arange(n_frames)[None] < lengths[:, None]).view(batch_size, 1, 1, n_frames).cuda()
    model_kwargs["y"]["scale"] = torch.ones(batch_size, device="cuda") * 2.5          # cgenerate.py:120
    sample = diffusion.p_sample_loop(model, (batch_size, model.njoints, model.nfeats, n_frames), clip_denoised=False,
                                     model_kwargs=model_kwargs, skip_timesteps=0, init_image=None, progress=True,
                                     dump_steps=None, noise=None, const_noise=False)
    sample_gf = gaussian_filter1d(sample.cpu().numpy(), sigma=1, axis=-1)
    sample = torch.from_numpy(sample_gf).to(sample.device)
    rot2xyz_pose_rep = 'xyz' if model.data_rep in ['xyz', 'hml_vec'] else model.data_rep
    rot2xyz_mask = None if rot2xyz_pose_rep == 'xyz' else model_kwargs['y']['mask'].reshape(batch_size, n_frames).bool()
    xyz = model.rot2xyz(x=sample, mask=rot2xyz_mask, pose_rep=rot2xyz_pose_rep, glob=True, translation=True,
                        jointstype="smplx", vertstrans=True, num_person=1, betas=None, beta=0, glob_rot=None,
                        get_rotations_back=False)
    assert tuple(xyz.shape) == (batch_size, 55, 3, n_frames) and torch.isfinite(xyz).all()
    # padded frames carry no joints: after re-rooting they hold exactly the (re-rooted) translation
    tr = sample[:, -1, :3]
    tr = tr - tr[:, :, [0]]
    pad = ~rot2xyz_mask
    assert torch.allclose(xyz.permute(0, 3, 1, 2)[pad], tr.permute(0, 2, 1)[pad][:, None, :].expand(-1, 55, -1), atol=1e-6)
    # and the masked frames equal the oracle's rot6d -> rotmat pushed through the same body model
    r = model.rot2xyz.rotations(sample, rot2xyz_mask)[0]["rotations"]
    want = sampler_ref.rotation_6d_to_matrix(sample.cpu()[:, :-1].permute(0, 3, 1, 2)[rot2xyz_mask.cpu()])
    assert (r.cpu() - want).abs().max() < 1e-5
