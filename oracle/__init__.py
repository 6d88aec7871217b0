"""CPU oracle for the ReGenNet diffusion-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``regennet_b200/`` may import this
package; the only permitted importers are ``tests/``, ``__graft_entry__.smoke``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

The oracle is a plain torch-fp32 / numpy-fp64 restatement of the reference
algorithm (liangxuy/ReGenNet) for

    SpacedDiffusion.p_sample_loop / ddim_sample_loop
      -> CMDM.forward(arch='online')
      -> posterior update
      -> rotation_6d_to_matrix

Every function cites the reference file:line it follows.  The transformer
arithmetic itself lives in a third-party dependency of the reference
(``torch.nn.TransformerDecoderLayer`` / ``MultiheadAttention``, pinned by the
reference at pytorch 1.7.1 / 1.12.0, not vendored); it is restated here from its
published post-norm algorithm.

PINNING: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the *imported reference itself*, generated in
the build container by ``tests/golden/make_golden.py`` (committed together with
the vectors it wrote under ``tests/golden/*.npz``); ``tests/test_oracle_golden.py``
checks the oracle against them on every CPU test run.
"""
