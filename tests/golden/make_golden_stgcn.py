"""Golden vectors for the evaluation feature extractor (ST-GCN) from the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden_stgcn.py

Writes tests/golden/stgcn.npz: the adjacency the reference's Graph builds for the file-free 'ntu-rgb+d' layout
(eval/a2m/recognition/models/stgcnutils/graph.py; the SMPL-X layout needs the licensed SMPLX_NEUTRAL.npz) and the reference's
`features` / `yhat` (eval/a2m/recognition/models/stgcn.py:76-126) for seeded synthetic weights and inputs, and prints the
oracle-vs-reference differences.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from oracle import ref_shim, stgcn_ref  # noqa: E402
import cases  # noqa: E402


def main():
    torch.set_num_threads(8)
    ref_shim.install()
    from eval.a2m.recognition.models.stgcn import STGCN
    from eval.a2m.recognition.models.stgcnutils import graph as ref_graph
    import tempfile
    # the 'smplx' layout reads kintree_table from SMPLX_NEUTRAL.npz (licensed, absent): point the UNMODIFIED reference at a
    # temporary .npz that holds the synthetic SMPL-X-shaped tree of cases.STGCN_SMPLX_PARENTS
    ktx = np.stack([np.array(cases.STGCN_SMPLX_PARENTS), np.arange(55)])
    with tempfile.NamedTemporaryFile(suffix=".npz", delete=False) as f:
        np.savez(f, kintree_table=ktx)
    ref_graph.SMPLX_KINTREE_PATH = f.name
    out = {}
    for name, c in cases.STGCN_CASES.items():
        model = STGCN(in_channels=c["in_channels"], num_class=c["num_class"], num_person=c["num_person"],
                      graph_args={"layout": c["layout"], "strategy": "spatial"}, edge_importance_weighting=True,
                      device="cpu")
        A = model.A.clone()
        sd = stgcn_ref.make_state_dict(A, c["in_channels"], c["num_class"], c["num_person"], seed=c["wseed"])
        model.load_state_dict(sd, strict=True)      # key names and shapes are the reference's
        model.eval()
        g = torch.Generator().manual_seed(c["xseed"])
        x = torch.randn(c["N"], A.shape[1], c["in_channels"], c["T"], generator=g)
        with torch.no_grad():
            ref = model({"output": x})
            feat, yhat = stgcn_ref.stgcn_forward(sd, x, c["num_person"])
        print("%-18s features absmax %.3f  oracle-vs-ref max abs: features %.3e  yhat %.3e" %
              (name, ref["features"].abs().max(), (ref["features"] - feat).abs().max(), (ref["yhat"] - yhat).abs().max()))
        out[name + ".A"] = A.numpy()
        out[name + ".features"] = ref["features"].numpy()
        out[name + ".yhat"] = ref["yhat"].numpy()
    # adjacency partitions of the reference's Graph for every file-free layout / strategy (float64, compared bit for bit)
    from eval.a2m.recognition.models.stgcnutils.graph import Graph
    for layout in cases.STGCN_GRAPH_LAYOUTS:
        for strategy in ("uniform", "distance", "spatial"):
            for hop in (1, 2):
                out["graph.%s.%s.%d" % (layout, strategy, hop)] = Graph(layout=layout, strategy=strategy, max_hop=hop).A
    # kinematic-tree layout ('smpl') through a synthetic tree written to a temporary kintree_table.pkl
    import pickle
    import tempfile
    kt = np.stack([np.array(cases.STGCN_SMPL_PARENTS), np.arange(24)])
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        pickle.dump(kt, f)
    out["graph.smpl.spatial.1"] = Graph(layout="smpl", strategy="spatial", kintree_path=f.name).A
    os.unlink(f.name)
    # evaluation metrics (eval/a2m/stgcn/{fid,diversity,accuracy}.py) on seeded feature matrices
    from eval.a2m.stgcn.accuracy import calculate_accuracy
    from eval.a2m.stgcn.diversity import calculate_diversity_multimodality
    from eval.a2m.stgcn.fid import calculate_fid
    import scipy.linalg as sla
    try:
        sla.sqrtm(np.eye(2), disp=False)
    except TypeError:   # this container's scipy (>= 1.18) dropped `disp`; give the unmodified reference the old signature
        _sqrtm = sla.sqrtm
        sla.sqrtm = lambda a, disp=True, **kw: _sqrtm(a, **kw) if disp else (_sqrtm(a, **kw), 0.0)
    f1, f2, labels = cases.metric_inputs()
    s1 = (np.mean(f1.numpy(), axis=0), np.cov(f1.numpy(), rowvar=False))
    s2 = (np.mean(f2.numpy(), axis=0), np.cov(f2.numpy(), rowvar=False))
    out["metrics.fid"] = np.float64(calculate_fid(s1, s2))
    out["metrics.fid_self"] = np.float64(calculate_fid(s1, s1))
    out["metrics.div_mm"] = np.array(calculate_diversity_multimodality(f1, labels, cases.METRIC_LABELS, seed=7))
    loader = [{"yhat": f1[i:i + 40, :cases.METRIC_LABELS], "y": labels[i:i + 40]} for i in range(0, f1.shape[0], 40)]
    acc, conf = calculate_accuracy(None, loader, cases.METRIC_LABELS, lambda b: b, "cpu")
    out["metrics.accuracy"] = np.float64(acc)
    out["metrics.confusion"] = conf.numpy()
    print("metrics: fid %.6f div/mm %s acc %.4f" % (out["metrics.fid"], out["metrics.div_mm"], acc))
    out["graph.smplx.spatial.1"] = Graph(layout="smplx", strategy="spatial").A
    os.unlink(ref_graph.SMPLX_KINTREE_PATH)
    np.savez(os.path.join(HERE, "stgcn.npz"), **out)


if __name__ == "__main__":
    main()
