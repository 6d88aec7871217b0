"""Golden-vector case table shared by make_golden.py (writer) and the tests (readers).

Model kwargs are exactly what the reference's utils/model_util.py:20-72 (get_model_args)
produces for the named dataset; the forward arithmetic only depends on
njoints/nfeats/cond_mode/num_actions/cm_mode.
"""

COMMON = dict(modeltype="", translation=True, pose_rep="rot6d", glob=True, glob_rot=True,
              latent_dim=512, ff_size=1024, num_layers=8, num_heads=4, dropout=0.1, activation="gelu",
              action_emb="tensor", arch="online", cm_mode="concat", body_model="smplx",
              wo_pos_emb=False, emb_trans_dec=False, clip_version="ViT-B/32")

MODELS = {
    # README.md:96 NTU120-AS online unconstrained (BASELINE.json configs 1, 2, 4)
    "ntu": dict(COMMON, njoints=56, nfeats=6, num_actions=26, num_frames=60, data_rep="rot6d",
                cond_mode="no_cond", cond_mask_prob=0.0, dataset="ntu"),
    # Chi3D action-conditioned with classifier-free guidance (config 3)
    "chi3d": dict(COMMON, njoints=56, nfeats=6, num_actions=8, num_frames=150, data_rep="rot6d",
                  cond_mode="action", cond_mask_prob=0.1, dataset="chi3d"),
    # HumanML-shaped text-conditioned (config 5); CLIP features injected
    "hml": dict(COMMON, njoints=263, nfeats=1, num_actions=1, num_frames=196, data_rep="hml_vec",
                cond_mode="text", cond_mask_prob=0.1, dataset="humanml"),
}


def synth_kw(name):
    m = MODELS[name]
    return dict(njoints=m["njoints"], nfeats=m["nfeats"], latent_dim=m["latent_dim"], ff_size=m["ff_size"],
                num_layers=m["num_layers"], cond_mode=m["cond_mode"], num_actions=m["num_actions"],
                clip_dim=512, cm_mode=m["cm_mode"])


# name -> dict(model, B, T, t (list) | loop spec)
FORWARD_CASES = {
    "fwd_ntu": dict(model="ntu", B=2, T=60, t=[999, 3], wseed=0, xseed=10),
    "fwd_ntu_b1_t0": dict(model="ntu", B=1, T=60, t=[0], wseed=1, xseed=11),
    "fwd_ntu_ragged_T37": dict(model="ntu", B=3, T=37, t=[5, 500, 77], wseed=0, xseed=12),
    "fwd_chi3d_cond": dict(model="chi3d", B=2, T=150, t=[640, 12], wseed=2, xseed=13),
    "fwd_chi3d_uncond": dict(model="chi3d", B=2, T=150, t=[640, 12], wseed=2, xseed=13, uncond=True),
    "fwd_chi3d_cfg": dict(model="chi3d", B=2, T=150, t=[640, 12], wseed=2, xseed=13, cfg_scale=2.5),
    "fwd_hml_text": dict(model="hml", B=2, T=196, t=[321, 900], wseed=3, xseed=14),
    # unconditional text model: mask_cond zeroes the CLIP features, embed_text still adds its bias
    "fwd_hml_uncond": dict(model="hml", B=2, T=196, t=[321, 900], wseed=3, xseed=14, uncond=True),
    "fwd_hml_cfg": dict(model="hml", B=1, T=50, t=[77], wseed=3, xseed=15, cfg_scale=2.5),
}

LOOP_CASES = {
    # ancestral sampling, 20 respaced steps (README.md:134-137 uses ddimK respacing with p_sample_loop)
    "loop_ntu_p20": dict(model="ntu", B=2, T=60, respacing="ddim20", ddim=False, wseed=0, xseed=10, seed=10),
    # full-length schedule, first 6 steps only is not expressible through the reference API; use 1000->'25'
    "loop_ntu_p25frac": dict(model="ntu", B=1, T=60, respacing="25", ddim=False, wseed=1, xseed=11, seed=3),
    "loop_chi3d_cfg_p10": dict(model="chi3d", B=2, T=150, respacing="ddim10", ddim=False, wseed=2, xseed=13,
                               seed=5, cfg_scale=2.5),
    "loop_hml_ddim10": dict(model="hml", B=2, T=196, respacing="ddim10", ddim=True, wseed=3, xseed=14,
                            seed=7, cfg_scale=2.5),
}

RESPACINGS = ["", "ddim5", "ddim20", "ddim100", "25", "10,15,20", "1000"]


# arch='offline' (model/cmdm.py:63-71, 228-238): nn.TransformerEncoder, condition token first, no causal mask.
# Separate golden file (forward_offline.npz, make_golden_offline.py) so the online goldens stay untouched.
OFFLINE_MODELS = {
    "ntu_off": dict(MODELS["ntu"], arch="offline"),
    "chi3d_off_add": dict(MODELS["chi3d"], arch="offline", cm_mode="add"),
    "hml_off": dict(MODELS["hml"], arch="offline"),
}


def synth_kw_offline(name):
    m = OFFLINE_MODELS[name]
    return dict(njoints=m["njoints"], nfeats=m["nfeats"], latent_dim=m["latent_dim"], ff_size=m["ff_size"],
                num_layers=m["num_layers"], cond_mode=m["cond_mode"], num_actions=m["num_actions"],
                clip_dim=512, cm_mode=m["cm_mode"], arch="offline")


OFFLINE_FORWARD_CASES = {
    "off_ntu": dict(model="ntu_off", B=2, T=60, t=[999, 3], wseed=4, xseed=20),
    "off_ntu_ragged_T37": dict(model="ntu_off", B=3, T=37, t=[5, 500, 77], wseed=4, xseed=21),
    "off_chi3d_add_cfg": dict(model="chi3d_off_add", B=2, T=150, t=[640, 12], wseed=5, xseed=22, cfg_scale=2.5),
    "off_hml_text": dict(model="hml_off", B=2, T=196, t=[321, 900], wseed=6, xseed=23),
}
OFFLINE_LOOP_CASES = {
    "off_loop_ntu_p10": dict(model="ntu_off", B=2, T=60, respacing="ddim10", ddim=False, wseed=4, xseed=20, seed=11),
}

# PLMS sampler (diffusion/gaussian_diffusion.py:1007-1202); goldens in loops_plms.npz (make_golden_plms.py).
# order 2 exercises the pseudo-improved-Euler first step (second model call at t - 1) and Adams-Bashforth 2;
# order 4 walks through Adams-Bashforth 2..4 as the history fills.  order=1 from a fresh loop raises TypeError in the
# reference (old_out is None at :1067), so it has no golden; the tests check that this package raises the same.
PLMS_LOOP_CASES = {
    "plms_ntu_o2": dict(model="ntu", B=2, T=60, respacing="ddim10", order=2, wseed=0, xseed=10, seed=21),
    "plms_ntu_o4": dict(model="ntu", B=1, T=37, respacing="ddim12", order=4, wseed=1, xseed=11, seed=22),
    "plms_chi3d_cfg_o3": dict(model="chi3d", B=2, T=150, respacing="ddim8", order=3, wseed=2, xseed=13, seed=23,
                              cfg_scale=2.5),
    "plms_ntu_o2_clip": dict(model="ntu", B=2, T=60, respacing="ddim6", order=2, wseed=0, xseed=10, seed=24, clip=True),
}

# arch='online' with cm_mode='add' (model/cmdm.py:207-211: x + cmotion embedding instead of fuse_process(concat)).
# Goldens in forward_add.npz / loops_add.npz (make_golden_add.py).
ADD_MODELS = {
    "ntu_add": dict(MODELS["ntu"], cm_mode="add"),
    "chi3d_add": dict(MODELS["chi3d"], cm_mode="add"),
}


def synth_kw_add(name):
    m = ADD_MODELS[name]
    return dict(njoints=m["njoints"], nfeats=m["nfeats"], latent_dim=m["latent_dim"], ff_size=m["ff_size"],
                num_layers=m["num_layers"], cond_mode=m["cond_mode"], num_actions=m["num_actions"],
                clip_dim=512, cm_mode="add")


ADD_FORWARD_CASES = {
    "add_ntu": dict(model="ntu_add", B=2, T=60, t=[999, 3], wseed=7, xseed=30),
    "add_ntu_ragged_T37": dict(model="ntu_add", B=3, T=37, t=[5, 500, 77], wseed=7, xseed=31),
    "add_chi3d_cfg": dict(model="chi3d_add", B=2, T=150, t=[640, 12], wseed=8, xseed=32, cfg_scale=2.5),
}
ADD_LOOP_CASES = {
    "add_loop_ntu_p10": dict(model="ntu_add", B=2, T=60, respacing="ddim10", ddim=False, wseed=7, xseed=30, seed=12),
}

# Evaluation feature extractor (ST-GCN, eval/a2m/recognition/models/stgcn.py); goldens in stgcn.npz (make_golden_stgcn.py).
# Oracle only so far (oracle/stgcn_ref.py): SURVEY.md 8f row 3 has no CUDA path yet.
STGCN_CASES = {
    # two persons (actor + reactor stacked along the feature axis, 6 rot6d features each), as eval/a2m/stgcn/evaluate.py:15-20
    "stgcn_p2": dict(layout="ntu-rgb+d", in_channels=12, num_class=26, num_person=2, N=3, T=60, wseed=0, xseed=40),
    # single person, odd length (the stride-2 blocks round 37 -> 19 -> 10)
    "stgcn_p1_T37": dict(layout="ntu-rgb+d", in_channels=6, num_class=8, num_person=1, N=2, T=37, wseed=1, xseed=41),
}
# the evaluation's real shape (eval/a2m/stgcn/evaluate.py:15-20: layout 'smplx', 56 nodes, two persons x 6 rot6d features):
# the reference reads the kinematic tree from the licensed SMPLX_NEUTRAL.npz; make_golden_stgcn.py hands it the tree below
# through a temporary .npz instead
STGCN_CASES["stgcn_smplx_p2"] = dict(layout="smplx", in_channels=12, num_class=26, num_person=2, N=3, T=60, wseed=2, xseed=42)
STGCN_GRAPH_LAYOUTS = ["ntu-rgb+d", "ntu_edge", "openpose"]
# a 55-joint kinematic tree shaped like SMPL-X's (22 body joints, jaw and two eyes on the head, 15 joints per hand in five
# 3-joint chains on each wrist); parent of joint j
STGCN_SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                       20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                       21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]


def stgcn_graph_args(c, ours):
    """graph_args of an STGCN case: this package takes the kinematic tree of the 'smplx' layout as ``kintree``, the reference
    reads it from a file (patched in by make_golden_stgcn.py)."""
    import numpy as np
    ga = {"layout": c["layout"], "strategy": "spatial"}
    if ours and c["layout"] == "smplx":
        ga["kintree"] = np.stack([np.array(STGCN_SMPLX_PARENTS), np.arange(55)])
    return ga

# the SMPL kinematic tree (parent of joint j), used to exercise the kintree-driven layouts without body-model files
STGCN_SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

METRIC_LABELS = 6


def metric_inputs():
    """Seeded feature matrices / labels for the evaluation-metric goldens (FID, diversity, multimodality, accuracy)."""
    import torch
    g = torch.Generator().manual_seed(123)
    f1 = torch.randn(120, 256, generator=g)
    f2 = 0.9 * torch.randn(150, 256, generator=g) + 0.1
    labels = torch.randint(0, METRIC_LABELS - 1, (120,), generator=g)   # the last label never occurs (zero quota)
    return f1, f2, labels


# model.rot2xyz(...) call sites (sample/cgenerate.py:156, eval/a2m/stgcn_eval.py:81): goldens in rot2xyz.npz
# (make_golden_rot2xyz.py) from the reference's Rotation2xyz / Rotation2xyz_x run with StubBodyModel in place of the
# smplx layers (their files are licensed and absent); everything around the skinning call is the reference's own code.
ROT2XYZ_CASES = {
    # name: body_model, B, J (incl. translation row), persons, T, masked tail frames per sample, glob, jointstype
    "x_p1_full": dict(body_model="smplx", B=3, J=56, P=1, T=9, cut=[0, 0, 0], glob=True, jointstype="smplx", seed=60),
    "x_p1_mask": dict(body_model="smplx", B=3, J=56, P=1, T=9, cut=[0, 4, 7], glob=True, jointstype="smplx", seed=61),
    "x_p2_mask": dict(body_model="smplx", B=2, J=56, P=2, T=7, cut=[2, 0], glob=True, jointstype="smplx", seed=62),
    "x_p2_vertices": dict(body_model="smplx", B=2, J=56, P=2, T=5, cut=[0, 1], glob=True, jointstype="vertices", seed=63),
    "x_p1_globrot": dict(body_model="smplx", B=2, J=56, P=1, T=6, cut=[1, 0], glob=False, jointstype="smplx", seed=64),
    "s_p1_mask": dict(body_model="smpl", B=3, J=25, P=1, T=8, cut=[0, 3, 5], glob=True, jointstype="smpl", seed=65),
    "s_p2_mask": dict(body_model="smpl", B=2, J=25, P=2, T=6, cut=[1, 0], glob=True, jointstype="vibe", seed=66),
    "s_p1_globrot": dict(body_model="smpl", B=2, J=25, P=1, T=5, cut=[0, 2], glob=False, jointstype="a2m", seed=67),
}
ROT2XYZ_GLOB_ROT = [3.141592653589793, 0.0, 0.0]   # the value the reference's callers use for glob=False


def rot2xyz_inputs(c):
    """Seeded pose tensor x [B, J, 6 P, T] and frame mask [B, T] (True = valid) of a ROT2XYZ case."""
    import torch
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.randn(c["B"], c["J"], 6 * c["P"], c["T"], generator=g)
    mask = torch.ones(c["B"], c["T"], dtype=torch.bool)
    for b, k in enumerate(c["cut"]):
        if k:
            mask[b, c["T"] - k:] = False
    return x, mask


def stub_body_model(kind):
    """Deterministic stand-in for model/smpl.py's SMPL / SMPLX layers with their call contract: ``num_betas``, keyword
    call, dict of [n, joints, 3] per jointstype.  Joints depend on every argument it is given (global orientation, each
    pose rotation, betas), so a wrong slice / order / mask in the caller changes the result."""
    import torch

    class StubBodyModel(torch.nn.Module):
        num_betas = 10

        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(7 if kind == "smplx" else 8)
            self.n_out = {"smplx": 55, "smpl": 24, "vibe": 49, "a2m": 18, "a2mpl": 30, "vertices": 40}
            self.register_buffer("rest", torch.randn(64, 3, generator=g))
            self.register_buffer("mix", 0.1 * torch.randn(64, 80, generator=g))
            self.register_buffer("dirs", torch.randn(80, 3, generator=g))
            self.register_buffer("shape_dirs", 0.05 * torch.randn(10, 64, 3, generator=g))

        def forward(self, body_pose=None, global_orient=None, betas=None, left_hand_pose=None, right_hand_pose=None,
                    return_verts=True, **kw):
            assert not kw, kw
            parts = [p for p in (body_pose, left_hand_pose, right_hand_pose) if p is not None]
            pose = torch.cat(parts, 1)                                        # [n, K, 3, 3]
            n, K = pose.shape[:2]
            local = torch.einsum("nkij,kj->nki", pose, self.dirs[:K])         # each rotation acts on its own direction
            pts = self.rest[None] + torch.einsum("jk,nki->nji", self.mix[:, :K], local)
            pts = pts + torch.einsum("nb,bji->nji", betas, self.shape_dirs)
            go = global_orient.reshape(n, 3, 3)
            pts = torch.einsum("nij,nkj->nki", go, pts)                       # [n, 64, 3]
            return {name: pts[:, :m] for name, m in self.n_out.items()}

    return StubBodyModel().eval()
