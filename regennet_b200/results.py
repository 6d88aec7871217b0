"""``results.npy`` writer of the sampling script (sample/cgenerate.py:167-192).

Host-side and format-only: the generated batches (already smoothed on the device by ``postprocess.gaussian_filter1d_time``)
are concatenated over repetitions, cut to ``num_samples * num_repetitions`` and written as the pickled dict the reference's
renderer (render/crendermotion.py) and evaluation scripts read back with ``np.load(..., allow_pickle=True).item()``:

    {'motion', 'output', 'cmotion', 'text', 'lengths', 'num_samples', 'num_repetitions'}

plus ``results.txt`` (one caption per line) and ``results_len.txt`` (one length per line).  Like the reference, an existing
output directory is REPLACED (:177-179).
"""
import os
import shutil

import numpy as np
import torch


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


class ResultsWriter:
    """Accumulates one entry per repetition, then ``save()``; mirrors the all_* lists of sample/cgenerate.py:102-166."""

    def __init__(self, num_samples, num_repetitions):
        self.num_samples = int(num_samples)
        self.num_repetitions = int(num_repetitions)
        self.all_motions, self.all_outputs, self.all_cmotions, self.all_lengths, self.all_text = [], [], [], [], []

    def add(self, motion, output, cmotion, lengths, text):
        """motion: rot2xyz result [bs, njoints, 3, T]; output: smoothed raw sample [bs, njoints, nfeats, T];
        cmotion: the actor's motion [bs, njoints, nfeats, T]; lengths [bs]; text: list of bs captions."""
        self.all_motions.append(_np(motion))
        self.all_outputs.append(_np(output))
        self.all_cmotions.append(_np(cmotion))
        self.all_lengths.append(_np(lengths))
        self.all_text += list(text)

    def save(self, out_path):
        """Write results.npy / results.txt / results_len.txt into ``out_path`` -> path of results.npy."""
        if not self.all_motions:
            raise ValueError("no repetitions were added")
        total = self.num_samples * self.num_repetitions
        motions = np.concatenate(self.all_motions, axis=0)[:total]
        outputs = np.concatenate(self.all_outputs, axis=0)[:total]
        cmotions = np.concatenate(self.all_cmotions, axis=0)[:total]
        text = self.all_text[:total]
        lengths = np.concatenate(self.all_lengths, axis=0)[:total]
        if os.path.exists(out_path):
            shutil.rmtree(out_path)
        os.makedirs(out_path)
        npy_path = os.path.join(out_path, 'results.npy')
        np.save(npy_path, {'motion': motions, 'output': outputs, 'cmotion': cmotions, 'text': text, 'lengths': lengths,
                           'num_samples': self.num_samples, 'num_repetitions': self.num_repetitions})
        with open(npy_path.replace('.npy', '.txt'), 'w') as fw:
            fw.write('\n'.join(text))
        with open(npy_path.replace('.npy', '_len.txt'), 'w') as fw:
            fw.write('\n'.join([str(l) for l in lengths]))
        return npy_path


def load_results(npy_path):
    """The dict written by ``ResultsWriter.save`` (what render/crendermotion.py reads)."""
    return np.load(npy_path, allow_pickle=True).item()
