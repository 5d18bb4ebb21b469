"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the RICK StyleGAN2 hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the CPU arm that is
timed next to the GPU path.  ``rick_b200`` never imports from here.

Contents
  ops_oracle.py      upfirdn2d / bias-act restatement (torch CPU)          -> op/upfirdn2d.py:159-200, op/fused_bias_act_kernel.cu:26-47
  model_oracle.py    functional G / D forward, losses, Fisher estimate     -> gan_training/models/model_probe_tune.py, train_dynamic_update_prune.py:82-118
  rick_oracle.py     Fisher accumulation -> per-filter FIM -> percentile
                     -> freeze/prune index sets -> mask application (NumPy) -> train_dynamic_update_prune.py:214-393, 427-437, 521-539
  csrc/oracle.c      plain-C restatement of the integer / order-sensitive
                     pieces (direct-form upfirdn2d, NumPy pairwise float32
                     mean, np.percentile 'linear', mask decisions)
  ref_loader.py      imports the *real* reference from /root/reference (only
                     exists in the authoring container) to pin the restatement
  make_golden.py     runs the real reference on CPU and writes tests/golden/*

Pinning status: the reference ships no golden vectors or tests,
so the oracle is pinned against outputs of the reference itself run in the
authoring container (tests/golden/*.npz, produced by make_golden.py).
"""
