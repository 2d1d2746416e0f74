"""CPU oracle for the on-policy rollout-to-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`xagents_b200/`) may import this
package; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do, and there only as the checker / the timed CPU baseline.

Parity status: PINNED for returns/GAE, n-step returns, the env-major flatten, minibatch
slicing, advantage normalisation and the loss forward pass -- `tests/golden/make_golden.py`
imports the reference's own `PPO` / `A2C` classes from `/root/reference` (under a NumPy-backed
shim of the TensorFlow ops they call, because TensorFlow is not installable in this image) and
executes the reference methods verbatim; the committed fixtures in `tests/golden/*.npz` are
those outputs and this oracle is checked against every one of them.  UNPINNED for the arithmetic
that lives inside un-vendored TensorFlow / tensorflow-probability (Categorical log-prob and
entropy, autodiff gradients): those follow the ops' published definitions
and are cross-checked against torch-CPU autograd in `oracle/torch_ref.py`.
"""
from .hotpath import *  # noqa: F401,F403
