"""Oracle (test infrastructure): the evaluation feature extractor ST-GCN, restated on CPU fp32.

Follows eval/a2m/recognition/models/stgcn.py:76-126 (STGCN.forward: person split, data_bn, ten st_gcn blocks, global
average pooling, person mean, 1x1 classifier), :145-213 (st_gcn block: graph convolution, BN-ReLU-temporal conv (9x1)-BN,
residual, ReLU) and eval/a2m/recognition/models/stgcnutils/tgcn.py:55-64 (1x1 conv to K*C_out channels, then
einsum('nkctv,kvw->nctw') with the K adjacency partitions).  Inference only (model.eval()): BatchNorm uses its running
statistics, dropout is the identity.  The convolution arithmetic belongs to torch (third party for the reference); it is
restated with matmul / explicit tap sums.

No CUDA path exists for this row yet: this file and its goldens (tests/golden/make_golden_stgcn.py) pin the oracle so that
the kernels can be built against it.  The adjacency `A` [K,V,V] is an input: the reference's SMPL-X layout reads its
kinematic tree from the licensed SMPLX_NEUTRAL.npz (stgcnutils/graph.py:74-81), which is not available; the goldens use the
file-free 'ntu-rgb+d' layout.
"""
import torch

# (in_channels, out_channels, temporal stride) of the ten blocks, stgcn.py:52-63
BLOCKS = [(None, 64, 1), (64, 64, 1), (64, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 1), (128, 128, 1),
          (128, 256, 2), (256, 256, 1), (256, 256, 1)]
BN_EPS = 1e-5


def _bn(x, sd, prefix, dim=1):
    """Inference BatchNorm over channel axis `dim`."""
    shape = [1] * x.dim()
    shape[dim] = -1
    mean, var = sd[prefix + "running_mean"].view(shape), sd[prefix + "running_var"].view(shape)
    w, b = sd[prefix + "weight"].view(shape), sd[prefix + "bias"].view(shape)
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


def _conv1x1(x, w, b, stride=1):
    """x [n,ci,t,v], w [co,ci,1,1] -> [n,co,ceil(t/stride),v]."""
    if stride > 1:
        x = x[:, :, ::stride]
    return torch.einsum("oc,nctv->notv", w[:, :, 0, 0], x) + b.view(1, -1, 1, 1)


def _tconv9(x, w, b, stride):
    """Temporal convolution, kernel (9,1), padding (4,0), stride (stride,1): x [n,c,t,v], w [co,ci,9,1]."""
    n, c, t, v = x.shape
    xp = torch.nn.functional.pad(x, (0, 0, 4, 4))
    t_out = (t + 8 - 9) // stride + 1
    out = torch.zeros(n, w.shape[0], t_out, v)
    for k in range(9):
        tap = xp[:, :, k:k + (t_out - 1) * stride + 1:stride]
        out = out + torch.einsum("oc,nctv->notv", w[:, :, k, 0], tap)
    return out + b.view(1, -1, 1, 1)


def st_gcn_block(sd, i, x, A, cin, cout, stride):
    p = "st_gcn_networks.%d." % i
    if i == 0:
        res = 0.0                                            # residual=False for the first block
    elif cin == cout and stride == 1:
        res = x
    else:
        res = _bn(_conv1x1(x, sd[p + "residual.0.weight"], sd[p + "residual.0.bias"], stride), sd, p + "residual.1.")
    K = A.shape[0]
    y = _conv1x1(x, sd[p + "gcn.conv.weight"], sd[p + "gcn.conv.bias"])          # [n, K*cout, t, v]
    n, kc, t, v = y.shape
    y = torch.einsum("nkctv,kvw->nctw", y.view(n, K, kc // K, t, v), A)
    y = torch.relu(_bn(y, sd, p + "tcn.0."))
    y = _bn(_tconv9(y, sd[p + "tcn.2.weight"], sd[p + "tcn.2.bias"], stride), sd, p + "tcn.3.")
    return torch.relu(y + res)


def stgcn_forward(sd, output, num_person):
    """output [N, V, C*num_person, T] (the sampler's `batch["output"]`) -> (features [N,256], yhat [N,num_class])."""
    if num_person == 2:
        N, V, C, T = output.shape
        C //= 2
        x = output.reshape(N, V, 2, C, T).permute(0, 3, 4, 1, 2)            # N, C, T, V, M
    else:
        x = output.permute(0, 2, 3, 1).unsqueeze(4)
    N, C, T, V, M = x.shape
    x = x.permute(0, 4, 3, 1, 2).contiguous()                                # N, M, V, C, T
    x = x.view(N, M * V * C, T) if num_person == 2 else x.view(N * M, V * C, T)
    x = _bn(x, sd, "data_bn.")
    x = x.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
    A = sd["A"]
    cin = C
    for i, (_, cout, stride) in enumerate(BLOCKS):
        x = st_gcn_block(sd, i, x, A * sd["edge_importance.%d" % i], cin, cout, stride)
        cin = cout
    feat = x.mean(dim=(2, 3)).view(N, M, -1).mean(dim=1)                     # avg_pool2d over (T,V), then persons
    yhat = feat @ sd["fcn.weight"][:, :, 0, 0].t() + sd["fcn.bias"]
    return feat, yhat


def make_state_dict(A, in_channels, num_class, num_person, seed=0):
    """Seeded synthetic weights with the reference's key names and shapes (no trained ST-GCN checkpoint exists in the
    reference tree).  BatchNorm statistics and affine parameters are randomised so that every term of the arithmetic
    is exercised; convolution weights are scaled like a trained network's (~1/sqrt(fan_in))."""
    g = torch.Generator().manual_seed(seed)
    K, V, _ = A.shape
    sd = {"A": A.clone().float()}

    def bn(prefix, c):
        sd[prefix + "weight"] = 1.0 + 0.2 * torch.randn(c, generator=g)
        sd[prefix + "bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + "running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[prefix + "running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + "num_batches_tracked"] = torch.tensor(100)

    def conv(prefix, co, ci, kt):
        sd[prefix + "weight"] = torch.randn(co, ci, kt, 1, generator=g) / (ci * kt) ** 0.5
        sd[prefix + "bias"] = 0.05 * torch.randn(co, generator=g)

    bn("data_bn.", in_channels * V)
    cin = in_channels // num_person
    for i, (_, cout, stride) in enumerate(BLOCKS):
        p = "st_gcn_networks.%d." % i
        conv(p + "gcn.conv.", cout * K, cin, 1)
        bn(p + "tcn.0.", cout)
        conv(p + "tcn.2.", cout, cout, 9)
        bn(p + "tcn.3.", cout)
        if i > 0 and not (cin == cout and stride == 1):
            conv(p + "residual.0.", cout, cin, 1)
            bn(p + "residual.1.", cout)
        sd["edge_importance.%d" % i] = 1.0 + 0.3 * torch.randn(K, V, V, generator=g)
        cin = cout
    conv("fcn.", num_class, 256, 1)
    return sd
