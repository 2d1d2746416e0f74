"""ACER's losses and output gradients against fixtures made by the REFERENCE'S OWN `ACER.update_gradients`
(tests/golden/make_golden.py --updates-only: the reference's code under the NumPy shim in float64, its tape answered by
central differences of the reference's own calculate_losses).  The arithmetic under test (agents/acer.py: calculate_losses,
calculate_grads) is device-agnostic torch, so this part runs on the host in float64; the same fixtures drive the whole
`update_gradients` on the GPU (Retrace kernel included) in tests/test_gpu_acer.py."""
import numpy as np
import pytest
import torch


def _bare_acer(g):
    from xagents_b200.agents.acer import ACER
    a = object.__new__(ACER)
    a.n_steps, a.n_envs, a.n_actions = int(g['n_steps']), int(g['n_envs']), int(g['n_actions'])
    a.epsilon, a.importance_c, a.delta = float(g['epsilon']), float(g['importance_c']), float(g['delta'])
    a.trust_region, a.entropy_coef, a.value_loss_coef = bool(g['trust_region']), float(g['entropy_coef']), float(g['value_loss_coef'])
    return a


@pytest.mark.parametrize('case', ['acer_update_trust_region', 'acer_update_plain'])
def test_acer_losses_and_output_gradients_vs_the_reference_run(golden, case):
    g = golden(case)
    a = _bare_acer(g)
    n = a.n_steps * a.n_envs
    t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)
    probs = t(g['action_probs']).requires_grad_(True)
    critic = t(g['full_critic_logits']).reshape(a.n_envs, a.n_steps + 1, a.n_actions)[:, :-1].reshape(n, a.n_actions).clone().requires_grad_(True)
    index = torch.as_tensor(g['actions'].astype(np.int64)).view(n, 1)
    sel_probs, sel_q = probs.gather(1, index).squeeze(1), critic.gather(1, index).squeeze(1)
    assert np.allclose(sel_q.detach().numpy(), g['selected_critic_logits'], rtol=0, atol=1e-12)
    values = t(g['values']).reshape(a.n_envs, a.n_steps + 1)[:, :-1].reshape(n)             # clip_last_step, env-major
    losses = a.calculate_losses(probs, values, t(g['returns']), sel_probs, t(g['selected_importance']), sel_q)
    if a.trust_region:
        assert abs(float(losses[0]) - float(g['loss'])) <= 1e-10 * abs(float(g['loss']))
        assert abs(float(losses[1]) - float(g['value_loss'])) <= 1e-10 * abs(float(g['value_loss']))
    else:
        assert abs(float(losses) - float(g['loss'])) <= 1e-10 * abs(float(g['loss']))
    d_probs, d_critic = a.calculate_grads(losses, probs, critic, t(g['avg_action_probs']))
    want_p = g['output_grads'] if a.trust_region else g['d_loss_d_action_probs']
    want_q = g['d_value_loss_d_critic_logits'] if a.trust_region else g['d_loss_d_critic_logits']
    # the fixture's gradients are central differences (h = 1e-6) of the reference's function: ~1e-8 of their scale
    assert np.abs(d_probs.numpy() - want_p).max() <= 1e-6 * np.abs(want_p).max()
    assert np.abs(d_critic.numpy() - want_q).max() <= 1e-6 * np.abs(want_q).max()
    if a.trust_region:                                             # the projection acted on some rows of this fixture
        free = -g['d_loss_d_action_probs'] / n
        assert (np.abs(free - want_p).max(-1) > 1e-6).sum() >= 5
