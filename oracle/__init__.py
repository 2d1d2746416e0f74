"""CPU oracle for the on-policy rollout-to-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`xagents_b200/`) may import this
package; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do, and there only as the checker / the timed CPU baseline.

Parity status: PINNED against the reference's own code.  `tests/golden/make_golden.py` imports the reference's
`PPO` / `A2C` / `ACER` / `TRPO` classes from `/root/reference` (under a NumPy-backed shim of the TensorFlow / tfp ops they
call, because TensorFlow is not installable in this image) and executes the reference methods verbatim; the committed
fixtures in `tests/golden/*.npz` are those outputs and this oracle is checked against every one of them:
returns/GAE, n-step returns, Retrace, the env-major flatten, minibatch slicing, advantage normalisation, the loss forward
passes (fp32), and -- round 2 -- the GRADIENTS: `grad_*.npz` hold central differences (float64 shim, h = 1e-6) of the loss that
the reference's own `PPO.update_gradients` / `A2C.train_step` computes, over every entry of the model outputs, for the three
distribution branches; `acer_update_*.npz` the reference's whole `ACER.update_gradients` (its tape answered the same way);
`trpo_losses.npz` the reference's `TRPO.calculate_losses` / `update_critic_weights`.  UNPINNED residue: the arithmetic INSIDE
un-vendored TensorFlow / tensorflow-probability ops (e.g. the summation order of `reduce_mean`, `Categorical.entropy`'s
`multiply_no_nan`) follows the ops' published definitions -- no TensorFlow runs here.
"""
from .hotpath import *  # noqa: F401,F403
