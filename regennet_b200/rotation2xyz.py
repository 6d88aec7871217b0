"""rot6d pose tensors -> rotation matrices (-> joints) -- mirror of the geometry half of the
reference's ``model/rotation2xyz.py``.

On the hot path: split persons on the feature axis, drop the
translation row, bring frames forward, select the masked frames and convert rot6d -> rotation
matrices with the library kernel (model/rotation2xyz.py:180-202, 253-270).  The SMPL / SMPL-X
linear-blend-skinning that follows (:217-240, 287-300) belongs to the third-party ``smplx`` package
and needs licensed body-model files; it is out of scope and only called when a ``smpl_model``
has been attached by the user.
"""
import torch

from .rotation_conversions import rotation_6d_to_matrix

JOINTSTYPES = ["a2m", "a2mpl", "smpl", "vibe", "vertices", "smplx"]


class Rotation2xyz:
    def __init__(self, device='cpu', dataset='amass', body_model='smplx'):
        self.device = device
        self.dataset = dataset
        self.body_model = body_model
        self.smpl_model = None  # attach smplx.SMPLXLayer / SMPLLayer here to get joints

    def rotations(self, x, mask=None, pose_rep="rot6d", translation=True, num_person=1):
        """x [B, J, F*num_person, T] -> list (one per person) of dicts
        {rotations [n, J-1(or J), 3, 3], translations [B, 3, T] or None} for the masked frames."""
        if pose_rep != "rot6d":
            raise NotImplementedError("only pose_rep='rot6d' is on the sampling hot path")
        if mask is None:
            mask = torch.ones((x.shape[0], x.shape[-1]), dtype=bool, device=x.device)
        out = []
        num_dim = x.shape[2] // num_person
        for xp in torch.split(x, num_dim, dim=2):
            if translation:
                x_translations = xp[:, -1, :3]
                x_rotations = xp[:, :-1]
            else:
                x_translations = None
                x_rotations = xp
            x_rotations = x_rotations.permute(0, 3, 1, 2)  # [B, T, J', 6]
            sel = x_rotations if bool(mask.all()) else x_rotations[mask]
            sel = sel.reshape(-1, x_rotations.shape[2], x_rotations.shape[3])
            out.append({"rotations": rotation_6d_to_matrix(sel), "translations": x_translations})
        return out

    def __call__(self, x, mask, pose_rep, translation, glob, jointstype, vertstrans, betas=None, beta=0,
                 glob_rot=None, num_person=1, **kwargs):
        if pose_rep == "xyz":
            return x
        if not glob and glob_rot is None:
            raise TypeError("You must specify global rotation if glob is False")
        if jointstype not in JOINTSTYPES:
            raise NotImplementedError("This jointstype is not implemented.")
        if self.smpl_model is None:
            raise RuntimeError("SMPL(-X) skinning is outside the B200 sampling hot path: the `smplx` package and its "
                               "licensed body-model files are not part of this build.  Use .rotations(x, mask, ...) for "
                               "the rot6d -> rotation-matrix stage or attach a body model as .smpl_model")
        raise NotImplementedError("joint regression through an attached body model is left to the reference's "
                                  "model/rotation2xyz.py; feed it the matrices from .rotations()")
