"""Golden vectors for the ``model.rot2xyz(...)`` call sites, written by running the UNMODIFIED
reference classes ``model/rotation2xyz.py::Rotation2xyz`` / ``Rotation2xyz_x`` on CPU (build container only):

    python tests/golden/make_golden_rot2xyz.py

The smplx body-model layers (licensed files, package absent) are replaced by cases.stub_body_model -- a deterministic
module with the same call contract -- so everything AROUND the skinning call is the reference's own code: person split,
translation row, mask select, rot6d -> rotmat, global orientation / pose slicing, scatter over the mask, re-rooting,
translation.  Inputs are re-created from seeds (cases.rot2xyz_inputs); only reference outputs are stored.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
import cases  # noqa: E402


def main():
    ref_shim.install()
    from model import rotation2xyz as R
    out = {}
    for name, c in cases.ROT2XYZ_CASES.items():
        cls = R.Rotation2xyz_x if c["body_model"] == "smplx" else R.Rotation2xyz
        r2x = object.__new__(cls)               # the constructor would load the licensed body-model files
        r2x.device, r2x.dataset = "cpu", "ntu"
        r2x.smpl_model = cases.stub_body_model(c["body_model"])
        x, mask = cases.rot2xyz_inputs(c)
        with contextlib.redirect_stdout(io.StringIO()):     # the smpl class prints shapes in its multi-person branch
            xyz = r2x(x=x, mask=mask, pose_rep="rot6d", glob=c["glob"], translation=True, jointstype=c["jointstype"],
                      vertstrans=True, num_person=c["P"], betas=None, beta=0,
                      glob_rot=None if c["glob"] else cases.ROT2XYZ_GLOB_ROT, get_rotations_back=False)
        out[name] = xyz.numpy()
        print("%-14s -> %s" % (name, tuple(xyz.shape)))
        if c["body_model"] == "smpl" and c["P"] == 1:
            _, rot, go = r2x(x=x, mask=mask, pose_rep="rot6d", glob=c["glob"], translation=True,
                             jointstype=c["jointstype"], vertstrans=True, num_person=1, betas=None, beta=0,
                             glob_rot=None if c["glob"] else cases.ROT2XYZ_GLOB_ROT, get_rotations_back=True)
            out[name + "/rotations"] = rot.numpy()
            out[name + "/global_orient"] = go.numpy()
    np.savez_compressed(os.path.join(HERE, "rot2xyz.npz"), **out)
    print("wrote rot2xyz.npz (%d arrays)" % len(out))


if __name__ == "__main__":
    main()
