"""fp64 restatement of ACER.update_gradients' arithmetic in the reference's own layout (env-major flat tensors, Python
Retrace loop, one tape) -- xagents/acer/agent.py:172-339.  Test infrastructure: the checker for agents/acer.py."""
import torch


def acer_output_gradients(probs_tm, q_tm, avg_probs_tm, rewards_tm, dones_tm, actions_tm, prev_probs_tm, *, gamma, epsilon,
                          importance_c, delta, trust_region, entropy_coef, value_loss_coef):
    """Inputs time-major: probs/q/avg_probs [(T+1), E, A]; rewards/dones/actions [T, E]; prev_probs [T, E, A].
    Returns (d_probs [(T+1),E,A], d_q [(T+1),E,A], returns [T,E]) in fp64, i.e. the gradients the reference's tape feeds
    into the network outputs (rows of the bootstrap step are zero)."""
    f64 = torch.float64
    T, E = rewards_tm.shape
    A = probs_tm.shape[-1]
    em = lambda x: x.to(f64).transpose(0, 1).contiguous()            # -> env-major [E, T(+1), ...]
    probs_full = em(probs_tm).reshape(E * (T + 1), A).clone().requires_grad_(True)
    q_full = em(q_tm).reshape(E * (T + 1), A).clone().requires_grad_(True)
    avg_full = em(avg_probs_tm).reshape(E * (T + 1), A)
    rewards, dones = em(rewards_tm).reshape(-1), em(dones_tm).reshape(-1)
    actions = em(actions_tm).reshape(-1).long()
    prev = em(prev_probs_tm).reshape(E * T, A)

    def clip_last_step(t):                                           # acer/agent.py:101-103
        return t.reshape((E, T + 1) + tuple(t.shape[1:]))[:, :T].reshape((E * T,) + tuple(t.shape[1:]))

    values = (probs_full * q_full).sum(-1)
    probs, avg, q = clip_last_step(probs_full), clip_last_step(avg_full), clip_last_step(q_full)
    rows = torch.arange(E * T, device=probs.device)
    sel_p, sel_q = probs[rows, actions], q[rows, actions]
    sel_imp = (probs / (prev + epsilon))[rows, actions]
    # calculate_returns (:186-208) on [E, T] slices
    imp_bar = torch.clamp(sel_imp, max=1.0).reshape(E, T)
    d, r, sq, v = dones.reshape(E, T), rewards.reshape(E, T), sel_q.reshape(E, T), values.reshape(E, T + 1)
    cur, rets = v[:, T], []
    for i in reversed(range(T)):
        cur = r[:, i] + gamma * cur * (1.0 - d[:, i])
        rets.append(cur)
        cur = imp_bar[:, i] * (cur - sq[:, i]) + v[:, i]
    returns = torch.stack(rets[::-1], 1).reshape(-1)
    # calculate_losses (:210-262)
    entropy = (-(probs * torch.log(probs + epsilon)).sum(1)).mean()
    adv = returns - clip_last_step(values)
    action_loss = -(torch.log(sel_p + epsilon) * (adv * torch.clamp(sel_imp, max=importance_c)).detach()).mean()
    value_loss = ((returns.detach() - sel_q) ** 2 * 0.5).mean() * value_loss_coef
    n = E * T
    if trust_region:                                                 # calculate_grads (:264-291)
        loss = -(action_loss - entropy_coef * entropy) * n
        g, = torch.autograd.grad(loss, [probs], retain_graph=True)
        k = -avg / (probs.detach() + epsilon)
        adj = torch.clamp(((k * g).sum(-1) - delta) / ((k ** 2).sum(-1) + epsilon), min=0.0)
        g = g - adj.reshape(n, 1) * k
        d_probs_full, = torch.autograd.grad(probs, [probs_full], grad_outputs=-g / n, retain_graph=True)
        d_q_full, = torch.autograd.grad(value_loss, [q_full])
    else:
        loss = action_loss + value_loss_coef * value_loss - entropy_coef * entropy
        d_probs_full, d_q_full = torch.autograd.grad(loss, [probs_full, q_full])
    tm = lambda x: x.reshape(E, T + 1, A).transpose(0, 1).contiguous()
    return tm(d_probs_full), tm(d_q_full), returns.reshape(E, T).transpose(0, 1).contiguous()
