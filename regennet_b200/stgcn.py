"""Evaluation feature extractor: ``STGCN`` with the reference's constructor, state-dict keys and ``forward(batch)`` contract
(eval/a2m/recognition/models/stgcn.py:29-126), inference only, computed by ``regen_stgcn_*`` (regennet_b200/csrc/stgcn.cu).

SURVEY.md 8f row 3.  The torch sub-modules below are parameter containers that give ``load_state_dict`` the reference's
key names and shapes (eval/a2m/stgcn/evaluate.py:24-26 loads a checkpoint into exactly this structure); none of them is
ever called -- the arithmetic runs in the library, and there is no CPU or PyTorch fallback.

STATUS: the CUDA path was written against the pinned oracle (oracle/stgcn_ref.py) after the round's GPU budget had been
spent; its GPU parity tests are marked xfail(strict=False) until they have been seen green on a B200.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib, stgcn_graph

_BLOCKS = [(None, 64, 1), (64, 64, 1), (64, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 1), (128, 128, 1),
           (128, 256, 2), (256, 256, 1), (256, 256, 1)]


class _GraphConv(nn.Module):   # key: gcn.conv.{weight,bias}   (stgcnutils/tgcn.py:46-53)
    def __init__(self, cin, cout, K):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout * K, kernel_size=(1, 1))


class _Block(nn.Module):       # keys of st_gcn (stgcn.py:171-203)
    def __init__(self, cin, cout, K, stride, residual=True):
        super().__init__()
        self.gcn = _GraphConv(cin, cout, K)
        self.tcn = nn.Sequential(nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                                 nn.Conv2d(cout, cout, (9, 1), (stride, 1), (4, 0)), nn.BatchNorm2d(cout),
                                 nn.Dropout(0.0, inplace=True))
        self.has_res_conv = residual and not (cin == cout and stride == 1)
        if self.has_res_conv:
            self.residual = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, stride=(stride, 1)), nn.BatchNorm2d(cout))


class _Handle:
    def __init__(self, ptr, device, key):
        self.ptr, self.device, self.key = ptr, device, key

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().regen_stgcn_destroy(self.ptr)
        except Exception:
            pass


class STGCN(nn.Module):
    """``STGCN(in_channels, num_class, num_person, graph_args, edge_importance_weighting, device, **kwargs)``.
    ``graph_args``: ``{"layout", "strategy"}`` as in the reference, plus ``"kintree"`` ([2, J] table) for the layouts whose
    edges the reference reads from body-model files ('smpl', 'smplx')."""

    def __init__(self, in_channels, num_class, num_person, graph_args, edge_importance_weighting, device, **kwargs):
        super().__init__()
        self.device = device
        self.in_channels = in_channels
        self.num_class = num_class
        self.num_person = num_person
        self.losses = ["accuracy", "cross_entropy", "mixed"]
        A = torch.tensor(stgcn_graph.adjacency(**graph_args), dtype=torch.float32, requires_grad=False)
        self.register_buffer('A', A)
        K, V = A.size(0), A.size(1)
        self.data_bn = nn.BatchNorm1d(in_channels * V)
        blocks, cin = [], in_channels // num_person
        for i, (_, cout, stride) in enumerate(_BLOCKS):
            blocks.append(_Block(cin, cout, K, stride, residual=i > 0))
            cin = cout
        self.st_gcn_networks = nn.ModuleList(blocks)
        if edge_importance_weighting:
            self.edge_importance = nn.ParameterList([nn.Parameter(torch.ones(A.size())) for _ in blocks])
        else:
            self.edge_importance = [1] * len(blocks)
        self.fcn = nn.Conv2d(256, num_class, kernel_size=1)
        self._handle = None

    # ---------------------------------------------------------------------------------------------
    def _packed(self, device):
        """The library's packed weight layout (include/regen_sm100.h)."""
        def bn(m):
            return [m.weight, m.bias, m.running_mean, m.running_var]
        parts = [self.A] + bn(self.data_bn)
        for i, blk in enumerate(self.st_gcn_networks):
            parts += [blk.gcn.conv.weight, blk.gcn.conv.bias] + bn(blk.tcn[0]) + [blk.tcn[2].weight, blk.tcn[2].bias] \
                + bn(blk.tcn[3])
            if blk.has_res_conv:
                parts += [blk.residual[0].weight, blk.residual[0].bias] + bn(blk.residual[1])
            imp = self.edge_importance[i]
            parts.append(imp if torch.is_tensor(imp) else torch.ones_like(self.A))
        parts += [self.fcn.weight, self.fcn.bias]
        return torch.cat([p.detach().to(device=device, dtype=torch.float32).reshape(-1) for p in parts]).contiguous()

    def _get_handle(self, device):
        key = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        h = self._handle
        if h is not None and h.device == device and h.key == key:
            return h
        if device.type != 'cuda':
            raise RuntimeError("regennet_b200.STGCN runs on CUDA (sm_100a) only; there is no CPU fallback")
        L = _lib.lib()
        desc = _lib.StgcnDesc(in_channels=self.in_channels, num_person=self.num_person, num_class=self.num_class,
                              num_node=self.A.size(1), num_part=self.A.size(0))
        packed = self._packed(device)
        if packed.numel() != L.regen_stgcn_packed_size(ctypes.byref(desc)):
            raise RuntimeError("packed ST-GCN weights: %d floats, the library expects %d"
                               % (packed.numel(), L.regen_stgcn_packed_size(ctypes.byref(desc))))
        hp = ctypes.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(L.regen_stgcn_create(ctypes.byref(hp), idx, ctypes.byref(desc)), "regen_stgcn_create")
        handle = _Handle(hp, device, key)
        _lib.check(L.regen_stgcn_load_weights(hp, _lib.ptr(packed), packed.numel(), _lib.stream_ptr(device)),
                   "regen_stgcn_load_weights")
        torch.cuda.current_stream(device).synchronize()   # `packed` is released below; the library keeps its own copy
        self._handle = handle
        return handle

    def forward(self, batch):
        """batch["output"]: [N, V, C * num_person, T] CUDA fp32 -> fills batch["features"] [N,256] and batch["yhat"]
        [N, num_class] (stgcn.py:76-126); inference mode only."""
        if self.training:
            raise NotImplementedError("STGCN training is outside the sampling / evaluation path; call .eval()")
        x = _lib.require_cuda_f32(batch["output"], 'batch["output"]').contiguous()
        N, V, C, T = x.shape
        if V != self.A.size(1) or C != self.in_channels:
            raise ValueError('batch["output"] must be [N,%d,%d,T], got %s' % (self.A.size(1), self.in_channels, tuple(x.shape)))
        h = self._get_handle(x.device)
        feat = torch.empty((N, 256), device=x.device, dtype=torch.float32)
        yhat = torch.empty((N, self.num_class), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().regen_stgcn_forward(h.ptr, _lib.ptr(x), N, T, _lib.ptr(feat), _lib.ptr(yhat),
                                                  _lib.stream_ptr(x.device)), "regen_stgcn_forward")
        batch["features"] = feat.squeeze()
        batch["yhat"] = yhat
        return batch

    def compute_accuracy(self, batch):
        """stgcn.py:128-136."""
        confusion = torch.zeros(self.num_class, self.num_class, dtype=int)
        yhat = batch["yhat"].max(dim=1).indices
        for label, pred in zip(batch["y"], yhat):
            confusion[label][pred] += 1
        return torch.trace(confusion) / torch.sum(confusion)
