"""Evaluation metrics over the ST-GCN features (host side): FID, diversity / multimodality, accuracy.

Same names, arguments and results as eval/a2m/stgcn/fid.py:6-61, eval/a2m/stgcn/diversity.py:6-79,
eval/a2m/stgcn/accuracy.py:4-14 and Evaluation.calculate_activation_statistics (eval/a2m/stgcn/evaluate.py:45-50).
These are a few hundred kiloflops on [N, 256] feature matrices (one 256 x 256 matrix square root per FID): host numpy /
scipy, like the reference; the feature extraction that feeds them is the library's part (regennet_b200/stgcn.py).
The random pair selection consumes numpy's global generator in the reference's order, so a seed gives the same numbers.
"""
import numpy as np
import torch
from scipy import linalg


def calculate_activation_statistics(activations):
    """-> (mean [D], covariance [D, D]) of an [N, D] feature matrix (torch tensor or array)."""
    a = activations.detach().cpu().numpy() if torch.is_tensor(activations) else np.asarray(activations)
    return np.mean(a, axis=0), np.cov(a, rowvar=False)


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """||mu1 - mu2||^2 + Tr(S1 + S2 - 2 sqrt(S1 S2)); the near-singular and complex-residue handling of fid.py:43-57."""
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape, 'Training and test mean vectors have different lengths'
    assert sigma1.shape == sigma2.shape, 'Training and test covariances have different dimensions'
    try:
        root, _ = linalg.sqrtm(sigma1.dot(sigma2), disp=False)
    except TypeError:   # scipy >= 1.18 dropped `disp` (the reference pins an older scipy); sqrtm then returns the root only
        root = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(root).all():
        print('fid calculation produces singular product; adding %s to diagonal of cov estimates' % eps)
        jitter = np.eye(sigma1.shape[0]) * eps
        root = linalg.sqrtm((sigma1 + jitter).dot(sigma2 + jitter))
    if np.iscomplexobj(root):
        if not np.allclose(np.diagonal(root).imag, 0, atol=1e-3):
            raise ValueError('Imaginary component {}'.format(np.max(np.abs(root.imag))))
        root = root.real
    delta = mu1 - mu2
    return delta.dot(delta) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(root)


def calculate_fid(statistics_1, statistics_2):
    return calculate_frechet_distance(statistics_1[0], statistics_1[1], statistics_2[0], statistics_2[1])


def calculate_diversity_multimodality(activations, labels, num_labels, seed=None, unconstrained=False):
    """Mean feature distance of 200 random pairs (diversity) and of 20 same-label pairs per label (multimodality)."""
    n_div, n_mm = 200, 20
    labels = labels.long()
    n = activations.shape[0]
    if seed is not None:
        np.random.seed(seed)
    first = np.random.randint(0, n, n_div)
    second = np.random.randint(0, n, n_div)
    diversity = 0
    for a, b in zip(first, second):
        diversity += torch.dist(activations[a, :], activations[b, :])
    diversity /= n_div

    quota = np.zeros(num_labels)
    quota[labels.unique()] = n_mm          # a label that never occurs keeps a zero quota
    multimodality = 0
    while np.any(quota > 0):
        a = np.random.randint(0, n)
        la = labels[a]
        if not quota[la]:
            continue
        b = np.random.randint(0, n)
        while labels[b] != la:
            b = np.random.randint(0, n)
        quota[la] -= 1
        multimodality += torch.dist(activations[a, :], activations[b, :])
    multimodality /= (n_mm * num_labels)
    return diversity.item(), multimodality.item()


def calculate_accuracy(model, motion_loader, num_labels, classifier, device):
    """-> (accuracy, confusion [num_labels, num_labels]); ``classifier(batch)["yhat"]`` are the ST-GCN logits."""
    confusion = torch.zeros(num_labels, num_labels, dtype=torch.long)
    with torch.no_grad():
        for batch in motion_loader:
            pred = classifier(batch)["yhat"].max(dim=1).indices
            for label, p in zip(batch["y"], pred):
                confusion[label][p] += 1
    return (torch.trace(confusion) / torch.sum(confusion)).item(), confusion
