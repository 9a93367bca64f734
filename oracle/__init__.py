"""CPU/torch ORACLE for the LoRA + textual-inversion training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sd_lora_trainer_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker / baseline, never as the thing shipped.

PARITY UNPINNED.  The reference (edenartlab/sd-lora-trainer) ships no tests,
golden vectors or known-answer fixtures for this path, and the
arithmetic lives in third-party packages that are absent from this image and
from /root/reference: diffusers==0.29.2, peft==0.10.0 (pyproject.toml:6,
requirements.txt:10).  This package therefore *restates* the published
behaviour of those packages next to a behaviour-exact
restatement of the reference-owned files (main.py:263-382, trainer/loss.py,
trainer/ti_cross_attn_loss.py, trainer/embedding_handler.py,
trainer/optimizer.py).  Golden vectors under tests/golden/ are produced by
this oracle itself (tests/golden/make_golden.py), which pins the oracle
against regressions but not against diffusers/peft.
"""
