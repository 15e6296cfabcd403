"""CPU oracle for the SegDistill distillation-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``segdistill_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker /
CPU baseline, never as the product path.

Parity status: the reference ships **no** golden vectors or known-answer tests
for this path.  The oracle is therefore pinned against the
reference itself: ``tests/golden/make_golden.py`` imports the unmodified
``/root/reference/mmseg/models/distillation/losses.py`` in the build container
and commits its outputs as fixtures under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement against those fixtures (and,
where ``/root/reference`` is mounted, against the live reference classes).
"""
from .kld_oracle import (  # noqa: F401
    kld_loss_torch, kld_closed_form_f64, mse_loss_torch, at_loss_torch,
    alpha_schedule, OracleKLD, ORACLE_PRESETS, make_preset, corr_loss_torch, ifvd_loss_torch,
)
from .seg_loss_oracle import decode_head_losses_torch  # noqa: F401
