"""Skeleton graph of the evaluation feature extractor (ST-GCN): adjacency partitions A [K, V, V].

Host-side mirror of eval/a2m/recognition/models/stgcnutils/graph.py (hop distances :147-160, column-normalised adjacency
:163-172, 'uniform' / 'distance' / 'spatial' partition strategies :108-144) for the layouts the reference's evaluation can
reach: 'ntu-rgb+d', 'ntu_edge', 'openpose' (edge lists spelled out in the reference, :40-104) and 'smpl' / 'smplx', whose
edges come from a kinematic tree.  The reference reads that tree from body-model files (SMPL kintree_table.pkl,
SMPLX_NEUTRAL.npz); here it is passed in as ``kintree`` (a [2, J] integer table: row 0 = parent, row 1 = joint id, like the
body-model files store it) so that no licensed file is needed inside this package.  SURVEY.md 8f row 3 groundwork: the
CUDA path of the extractor is not built yet; this module and oracle/stgcn_ref.py are pinned against the reference.
"""
import numpy as np

_NTU_1BASE = [(1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9), (11, 10), (12, 11),
              (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (22, 23), (23, 8), (24, 25),
              (25, 12)]
_NTU_EDGE_1BASE = [(1, 2), (3, 2), (4, 3), (5, 2), (6, 5), (7, 6), (8, 7), (9, 2), (10, 9), (11, 10), (12, 11), (13, 1),
                   (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (21, 22), (22, 8), (23, 24),
                   (24, 12)]
_OPENPOSE = [(4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5), (8, 2), (5, 1), (2, 1), (0, 1),
             (15, 0), (14, 0), (17, 15), (16, 14)]


def _edges(layout, kintree):
    """-> (num_node, neighbour links, centre joint)."""
    if layout == 'openpose':
        return 18, list(_OPENPOSE), 1
    if layout == 'ntu-rgb+d':
        return 25, [(i - 1, j - 1) for i, j in _NTU_1BASE], 21 - 1
    if layout == 'ntu_edge':
        return 24, [(i - 1, j - 1) for i, j in _NTU_EDGE_1BASE], 2
    if layout in ('smpl', 'smplx'):
        if kintree is None:
            raise ValueError("layout %r needs the body model's kinematic tree (kintree=[2, J] table)" % layout)
        kt = np.asarray(kintree)
        joints = 24 if layout == 'smpl' else 55
        link = [(int(k), int(kt[1][i + 1])) for i, k in enumerate(kt[0][1:])]
        link.append((0, joints))          # root rotation <-> root translation (the extra node)
        return joints + 1, link, 0
    raise NotImplementedError("This Layout is not supported")


def hop_distance(num_node, edge, max_hop=1):
    A = np.zeros((num_node, num_node))
    for i, j in edge:
        A[j, i] = 1
        A[i, j] = 1
    hop = np.zeros((num_node, num_node)) + np.inf
    reach = np.stack([np.linalg.matrix_power(A, d) for d in range(max_hop + 1)]) > 0
    for d in range(max_hop, -1, -1):
        hop[reach[d]] = d
    return hop


def normalize_digraph(A):
    deg = np.sum(A, 0)
    Dn = np.zeros_like(A)
    for i in range(A.shape[0]):
        if deg[i] > 0:
            Dn[i, i] = deg[i] ** (-1)
    return np.dot(A, Dn)


def adjacency(layout='ntu-rgb+d', strategy='spatial', max_hop=1, dilation=1, kintree=None):
    """float64 [K, V, V], identical to ``Graph(layout, strategy, ...).A`` of the reference."""
    V, link, center = _edges(layout, kintree)
    edge = [(i, i) for i in range(V)] + link
    hop = hop_distance(V, edge, max_hop)
    valid = range(0, max_hop + 1, dilation)
    adj = np.zeros((V, V))
    for h in valid:
        adj[hop == h] = 1
    norm = normalize_digraph(adj)
    if strategy == 'uniform':
        return norm[None].copy()
    if strategy == 'distance':
        A = np.zeros((len(valid), V, V))
        for i, h in enumerate(valid):
            A[i][hop == h] = norm[hop == h]
        return A
    if strategy == 'spatial':
        parts = []
        for h in valid:
            root, close, further = np.zeros((V, V)), np.zeros((V, V)), np.zeros((V, V))
            for i in range(V):
                for j in range(V):
                    if hop[j, i] == h:
                        if hop[j, center] == hop[i, center]:
                            root[j, i] = norm[j, i]
                        elif hop[j, center] > hop[i, center]:
                            close[j, i] = norm[j, i]
                        else:
                            further[j, i] = norm[j, i]
            if h == 0:
                parts.append(root)
            else:
                parts.append(root + close)
                parts.append(further)
        return np.stack(parts)
    raise NotImplementedError("This Strategy is not supported")
