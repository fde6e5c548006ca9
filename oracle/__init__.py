"""CPU oracle for the mae_clip training-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mae_clip_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the timed CPU baseline - never as the product path.

Parity status
-------------
* L1-L7 (ProjectionHead, logits, soft targets, bidirectional CE, backward):
  PINNED - ``tests/golden/make_golden.py`` imports the unmodified reference
  (``/root/reference/CLIP.py``, ``modules.py``) in the dev container and the
  committed fixtures under ``tests/golden/`` hold its outputs; the oracle is
  checked against every fixture by ``tests/test_oracle_golden.py``.
* M1-M3 (MAE random masking, patchify + norm-pix target, masked MSE):
  PARITY UNPINNED - the reference tree contains no MAE code at all
  (SURVEY.md section 0.2), so ``oracle/mae_ref.py`` is the *defining*
  restatement of the published MAE formulation, not a check against
  reference outputs.
"""
