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


GRAD_CASES = ['grad_ppo_logits', 'grad_ppo_probs', 'grad_ppo_normal', 'grad_a2c_logits', 'grad_a2c_probs', 'grad_a2c_normal']


@pytest.mark.parametrize('case', GRAD_CASES)
def test_oracle_loss_gradients_vs_differences_of_the_reference_loss(golden, case):
    """The closed-form d loss / d(actor_output, critic_output) of the oracle (what the CUDA loss kernels are compared with)
    against central differences of the loss computed by the REFERENCE'S OWN PPO.update_gradients / A2C.train_step under the
    float64 shim (make_golden.py --gradients-only): the gradients are pinned to the reference, not to a restatement."""
    import oracle
    g = golden(case)
    kind, actor_kind = str(g['kind']), str(g['actor_kind'])
    f64 = np.float64
    if actor_kind == 'normal':
        pytest.skip('the oracle keeps closed forms for the categorical heads; the normal head is checked on the device')
    args = (g['actor_output'], g['critic_output'], g['actions'], g['old_values'], g['returns'])
    if kind == 'ppo':
        d_actor, d_values = oracle.ppo_loss_grads(*args, g['old_log_probs'], g['advantages'], float(g['clip_norm']), float(g['entropy_coef']),
                                                  float(g['value_loss_coef']), actor_kind == 'probs', f64)
    else:
        d_actor, d_values = oracle.a2c_loss_grads(*args, float(g['entropy_coef']), float(g['value_loss_coef']), actor_kind == 'probs', f64)
    assert np.abs(d_actor - g['d_actor']).max() <= 1e-6 * np.abs(g['d_actor']).max()
    assert np.abs(d_values - g['d_critic']).max() <= 1e-6 * np.abs(g['d_critic']).max()
