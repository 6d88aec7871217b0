"""rot6d pose tensors -> rotation matrices -> joints: mirror of the reference's ``model/rotation2xyz.py``
(``Rotation2xyz`` for ``body_model='smpl'``, ``Rotation2xyz_x`` for ``'smplx'``; constructed by ``CMDM.__init__``,
model/cmdm.py:107-111, and called by sample/cgenerate.py:156 and eval/a2m/stgcn_eval.py:81 right after sampling).

On the hot path: split persons on the feature axis, drop the translation row, bring
frames forward, select the masked frames and convert rot6d -> rotation matrices with the library kernel
(model/rotation2xyz.py:35-56, 186-202, 253-270).  The SMPL / SMPL-X linear-blend skinning that follows
(:68-70, 217-224, 287-294) belongs to the third-party ``smplx`` package and its licensed body-model files: it is
delegated to the body-model layer in ``self.smpl_model`` exactly as the reference does, and everything after it (joint
selection, scatter back over the mask, re-rooting, translation) is restated here.  ``smpl_model`` is built like the
reference's ``model/smpl.py`` wrappers when ``smplx`` and the model files are available; otherwise it stays ``None``
until the user attaches one (``attach_body_model``), and only then does ``__call__`` raise.

The two reference classes differ in small ways that callers can observe; both behaviours are kept (selected by
``body_model``): SMPL passes ``body_pose`` = all non-root rotations, SMPL-X splits body / hands (jaw and eyes are
dropped, :207-222); SMPL always drops rotation 0 from the pose even with ``glob=False`` (:66), SMPL-X only when it is
the global orientation (:200-205); with ``num_person > 1`` the translations are added un-rooted (:85-89, 230-233), with
one person they are re-rooted at frame 0 (:150-154, 306-310); only ``Rotation2xyz`` honours ``get_rotations_back``.
"""
import os

import torch

from . import rotation_conversions as geometry

JOINTSTYPES = ["a2m", "a2mpl", "smpl", "vibe", "smplx", "vertices"]
# model/smpl.py:19-23
JOINTSTYPE_ROOT = {"a2m": 0, "smpl": 0, "smplx": 0, "a2mpl": 0, "vibe": 8}


def _default_body_model(body_model):
    """The reference builds ``SMPL()`` / ``SMPLX()`` (model/smpl.py:65-113: smplx layers + joint maps) from files under
    ./body_models.  Neither ``smplx`` nor the licensed files ship with this package; when both are present the same
    layers are built here, otherwise None."""
    try:
        import numpy as np
        import smplx  # noqa: F401
    except Exception:
        return None
    path = os.environ.get("REGEN_BODY_MODEL_PATH", "./body_models")
    if not os.path.isdir(os.path.join(path, "smplx" if body_model == 'smplx' else "smpl")):
        return None
    try:
        if body_model == 'smplx':
            from smplx import SMPLXLayer

            class SMPLX(SMPLXLayer):
                def __init__(self, **kw):
                    super().__init__(model_path=os.path.join(path, "smplx"), **kw)
                    self.maps = {"smplx": np.arange(55)}

                def forward(self, *a, **k):
                    o = super().forward(*a, **k)
                    out = {"vertices": o.vertices}
                    for name, idx in self.maps.items():
                        out[name] = o.joints[:, idx]
                    return out
            return SMPLX().eval()
        from smplx import SMPLLayer

        class SMPL(SMPLLayer):
            def __init__(self, **kw):
                super().__init__(model_path=os.path.join(path, "smpl"), **kw)
                self.maps = {"smpl": np.arange(24)}

            def forward(self, *a, **k):
                o = super().forward(*a, **k)
                out = {"vertices": o.vertices}
                for name, idx in self.maps.items():
                    out[name] = o.joints[:, idx]
                return out
        return SMPL().eval()
    except Exception:
        return None


class Rotation2xyz:
    def __init__(self, device='cpu', dataset='amass', body_model='smplx'):
        self.device = device
        self.dataset = dataset
        self.body_model = body_model
        self.smpl_model = _default_body_model(body_model)
        if self.smpl_model is not None:
            self.smpl_model = self.smpl_model.to(device)

    def attach_body_model(self, layer):
        """Use `layer` for the skinning step.  Contract (model/smpl.py): ``layer.num_betas``; called as
        ``layer(body_pose=..., global_orient=..., betas=...)`` (smpl) or ``layer(betas=..., body_pose=..., left_hand_pose=...,
        right_hand_pose=..., global_orient=..., return_verts=True)`` (smplx); returns a dict with a ``[n, joints, 3]`` entry
        per jointstype."""
        self.smpl_model = layer
        return self

    # ------------------------------------------------------------------------------------------ rotations (GPU)
    @staticmethod
    def _person_rotations(xp, mask, all_frames, pose_rep, translation):
        """One person's slice xp [B, J, F, T] -> (rotations [n, J', 3, 3] of the masked frames, translations [B, 3, T] or
        None, (B, T)).  model/rotation2xyz.py:35-56."""
        if translation:
            x_translations = xp[:, -1, :3]
            x_rotations = xp[:, :-1]
        else:
            x_translations = None
            x_rotations = xp
        x_rotations = x_rotations.permute(0, 3, 1, 2)                     # [B, T, J', F]
        nsamples, time, njoints, feats = x_rotations.shape
        sel = x_rotations.reshape(nsamples * time, njoints, feats) if all_frames else x_rotations[mask]
        if pose_rep == "rot6d":
            rotations = geometry.rotation_6d_to_matrix(sel)                # library kernel (60 B per rotation)
        elif pose_rep == "rotvec":
            rotations = geometry.axis_angle_to_matrix(sel)
        elif pose_rep == "rotmat":
            rotations = sel.reshape(-1, njoints, 3, 3)
        elif pose_rep == "rotquat":
            rotations = geometry.quaternion_to_matrix(sel)
        else:
            raise NotImplementedError("No geometry for this one.")
        return rotations, x_translations, (nsamples, time)

    def rotations(self, x, mask=None, pose_rep="rot6d", translation=True, num_person=1):
        """x [B, J, F*num_person, T] -> list (one per person) of dicts
        {rotations [n, J-1 (or J), 3, 3] of the masked frames, translations [B, 3, T] or None}."""
        if mask is None:
            mask = torch.ones((x.shape[0], x.shape[-1]), dtype=bool, device=x.device)
        mask = mask.to(x.device)
        all_frames = bool(mask.all())
        out = []
        num_dim = x.shape[2] // num_person
        for xp in torch.split(x, num_dim, dim=2):
            rot, tr, _ = self._person_rotations(xp, mask, all_frames, pose_rep, translation)
            out.append({"rotations": rot, "translations": tr})
        return out

    # ------------------------------------------------------------------------------------------ joints
    def _joints(self, rotations, glob, glob_rot, betas, beta, jointstype, multi_person):
        """rotation matrices of the masked frames -> (joints [n, Jout, 3], rotations, global_orient, betas) through the
        attached body model (model/rotation2xyz.py:58-72 smpl, :196-224 / 279-294 smplx)."""
        smplx_like = self.body_model == 'smplx'
        if not glob:
            global_orient = torch.tensor(glob_rot, device=rotations.device)
            global_orient = geometry.axis_angle_to_matrix(global_orient).view(1, 1, 3, 3)
            global_orient = global_orient.repeat(len(rotations), 1, 1, 1)
            if not smplx_like:
                rotations = rotations[:, 1:]
        else:
            global_orient = rotations[:, 0, None] if (smplx_like and multi_person) else rotations[:, 0]
            rotations = rotations[:, 1:]
        if betas is None:
            betas = torch.zeros([rotations.shape[0], self.smpl_model.num_betas], dtype=rotations.dtype,
                                device=rotations.device)
            betas[:, 1] = beta
        if smplx_like:
            out = self.smpl_model(betas=betas, body_pose=rotations[:, 0:21], left_hand_pose=rotations[:, 24:39],
                                  right_hand_pose=rotations[:, 39:54], global_orient=global_orient, return_verts=True)
        else:
            out = self.smpl_model(body_pose=rotations, global_orient=global_orient, betas=betas)
        return out[jointstype], rotations, global_orient, betas

    def __call__(self, x, mask, pose_rep, translation, glob, jointstype, vertstrans, betas=None, beta=0,
                 glob_rot=None, num_person=1, get_rotations_back=False, **kwargs):
        if pose_rep == "xyz":
            return x
        if mask is None:
            mask = torch.ones((x.shape[0], x.shape[-1]), dtype=bool, device=x.device)
        if not glob and glob_rot is None:
            raise TypeError("You must specify global rotation if glob is False")
        if jointstype not in JOINTSTYPES:
            raise NotImplementedError("This jointstype is not implemented.")
        if self.smpl_model is None:
            raise RuntimeError("Rotation2xyz needs a body model for the skinning step: the `smplx` package and its "
                               "licensed model files (./body_models, or REGEN_BODY_MODEL_PATH) were not found.  Attach "
                               "one with model.rot2xyz.attach_body_model(layer), or use .rotations(x, mask, ...) for "
                               "the rot6d -> rotation-matrix stage alone")
        mask = mask.to(x.device)
        all_frames = bool(mask.all())
        multi = num_person > 1
        num_dim = x.shape[2] // num_person if multi else x.shape[2]
        parts = []
        rotations = global_orient = None
        for xp in (torch.split(x, num_dim, dim=2) if multi else (x,)):
            rotations, x_translations, (nsamples, time) = self._person_rotations(xp, mask, all_frames, pose_rep,
                                                                                 translation)
            # betas are created for the first person and re-used for the others, as in the reference loop (:68-71)
            joints, rotations, global_orient, betas = self._joints(rotations, glob, glob_rot, betas, beta, jointstype,
                                                                   multi)
            if all_frames:
                x_xyz = joints.reshape(nsamples, time, joints.shape[1], 3).to(xp.dtype)
            else:
                x_xyz = torch.empty(nsamples, time, joints.shape[1], 3, device=xp.device, dtype=xp.dtype)
                x_xyz[~mask] = 0
                x_xyz[mask] = joints
            x_xyz = x_xyz.permute(0, 2, 3, 1).contiguous()                # [B, Jout, 3, T]
            if jointstype != "vertices":                                  # root joint at the origin
                rootindex = JOINTSTYPE_ROOT[jointstype]
                x_xyz = x_xyz - x_xyz[:, [rootindex], :, :]
            if translation and vertstrans:
                if not multi:                                             # one person: first frame at the origin
                    x_translations = x_translations - x_translations[:, :, [0]]
                x_xyz = x_xyz + x_translations[:, None, :, :]
            parts.append(x_xyz)
        x_xyz = torch.cat(parts, 2) if multi else parts[0]
        if get_rotations_back and self.body_model != 'smplx':
            return x_xyz, rotations, global_orient
        return x_xyz
