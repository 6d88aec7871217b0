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
    np.savez(os.path.join(HERE, "stgcn.npz"), **out)


if __name__ == "__main__":
    main()
