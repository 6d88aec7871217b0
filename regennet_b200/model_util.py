"""Model / diffusion factory with the call surface of the reference's ``utils/model_util.py``:
``create_model_and_diffusion(args, data)``, ``get_model_args``, ``create_gaussian_diffusion``, ``load_model_wo_clip``.

``args`` is the reference's argparse namespace (utils/parser_util.py), ``data`` its DataLoader (only
``data.dataset.num_actions`` / ``num_person`` are read).
"""
from . import gaussian_diffusion as gd
from .cmdm import CMDM
from .respace import SpacedDiffusion, space_timesteps

# skeleton size per body model and feature width per pose representation (utils/model_util.py:38-47)
_JOINTS = {"smpl": 25, "smplx": 56}
_FEATS = {"rot6d": 6, "xyz": 3}
# text datasets use the HumanML vector representation instead (:49-56): (data_rep, njoints, nfeats)
_HML = {"humanml": ("hml_vec", 263, 1), "kit": ("hml_vec", 251, 1)}
_FRAMES = {"ntu": 60, "chi3d": 150}
_DIFFUSION_STEPS = 1000


def load_model_wo_clip(model, state_dict):
    """Load a checkpoint that was saved without the frozen CLIP weights: nothing unexpected, only clip_model.* missing."""
    result = model.load_state_dict(state_dict, strict=False)
    assert len(result.unexpected_keys) == 0
    assert all([name.startswith('clip_model.') for name in result.missing_keys])


def create_model_and_diffusion(args, data):
    if args.setting == 'cmdm':
        model = CMDM(**get_model_args(args, data))
        args.num_person = 1   # the reactor alone is diffused; the actor is the condition (utils/model_util.py:15)
    return model, create_gaussian_diffusion(args)


def _cond_mode(args):
    if args.unconstrained:
        return 'no_cond'
    return 'text' if args.dataset in _HML else 'action'


def get_model_args(args, data):
    """Constructor kwargs of CMDM for the reference's command-line arguments (utils/model_util.py:20-72)."""
    ds = data.dataset
    body_model = args.body_model
    data_rep = args.pose_rep
    shape = {}
    if body_model in _JOINTS:
        shape['njoints'] = _JOINTS[body_model]
    if data_rep in _FEATS:
        shape['nfeats'] = _FEATS[data_rep]
    if args.dataset in _HML:
        data_rep, shape['njoints'], shape['nfeats'] = _HML[args.dataset]
    kw = dict(modeltype='', translation=True, pose_rep='rot6d', glob=True, glob_rot=True, ff_size=1024, num_heads=4,
              dropout=0.1, activation="gelu", action_emb='tensor', clip_version='ViT-B/32')
    kw.update(njoints=shape['njoints'], nfeats=shape['nfeats'],
              num_actions=getattr(ds, 'num_actions', 1), num_person=getattr(ds, 'num_person', 1),
              num_frames=_FRAMES[args.dataset], latent_dim=args.latent_dim, num_layers=args.layers,
              data_rep=data_rep, cond_mode=_cond_mode(args), cond_mask_prob=args.cond_mask_prob, arch=args.arch,
              cm_mode=args.cm_mode, body_model=body_model, wo_pos_emb=args.wo_pos_emb,
              emb_trans_dec=args.emb_trans_dec, dataset=args.dataset)
    return kw


def create_gaussian_diffusion(args):
    """The sampler the reference builds (utils/model_util.py:75-117): 1000 base steps of the named schedule, x0
    prediction, fixed variance (small or large), optional respacing."""
    respacing = args.timestep_respacing or [_DIFFUSION_STEPS]
    var_type = gd.ModelVarType.FIXED_SMALL if args.sigma_small else gd.ModelVarType.FIXED_LARGE
    loss_weights = {name: getattr(args, name) for name in
                    ('lambda_vel', 'lambda_rcxyz', 'lambda_fc', 'lambda_orient', 'lambda_body', 'lambda_transl')}
    return SpacedDiffusion(
        use_timesteps=space_timesteps(_DIFFUSION_STEPS, respacing),
        betas=gd.get_named_beta_schedule(args.noise_schedule, _DIFFUSION_STEPS, 1.),
        model_mean_type=gd.ModelMeanType.START_X,
        model_var_type=var_type,
        loss_type=gd.LossType.MSE,
        rescale_timesteps=False,
        data_rep=args.pose_rep,
        num_person=args.num_person,
        body_model=args.body_model,
        vel_threshold=args.vel_threshold,
        **loss_weights,
    )
