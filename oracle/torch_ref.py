"""torch-CPU stand-ins for the TensorFlow ops on the path (checker / CPU-baseline timing only).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  TensorFlow is not installable in this
image, so `tf.gather`, the reductions and `tf.GradientTape` of xagents/ppo/agent.py:96-137,149-154
are stood in for by the torch-CPU ops with the same published semantics.  Autograd here plays the
part of the tape: it gives gradients that are independent of the closed forms in hotpath.py.
torch's `maximum` splits tied gradients 50/50 whereas TF sends them to the first argument, so the
losses are written with explicit `where(a >= b, a, b)` selections, which is TF's rule.
"""
import torch


def _tf_max(a, b):
    return torch.where(a >= b, a, b)


def _tf_clip(x, lo, hi):
    # gradient 1 on the closed interval lo <= x <= hi, else 0 (tf.clip_by_value)
    return torch.where(x < lo, torch.full_like(x, lo), torch.where(x > hi, torch.full_like(x, hi), x))


def categorical(actor_output, actions, is_probs=False):
    logits = torch.log(actor_output) if is_probs else actor_output
    lsm = torch.log_softmax(logits, dim=-1)
    logp = lsm.gather(-1, actions.long().view(-1, 1)).squeeze(-1)
    ent = -(lsm.exp() * lsm).sum(-1)
    return logp, ent


def ppo_loss_autograd(actor_output, values, actions, old_values, returns, old_log_probs, advantages,
                      clip_norm, entropy_coef, value_loss_coef, is_probs=False, dtype=torch.float32):
    """xagents/ppo/agent.py:112-134 with autograd as the tape. Returns (scalars dict, d_actor, d_values)."""
    t = lambda x: torch.as_tensor(x, dtype=dtype)
    actor_output = t(actor_output).clone().requires_grad_(True)
    values = t(values).clone().requires_grad_(True)
    old_values, returns, old_log_probs, advantages = t(old_values), t(returns), t(old_log_probs), t(advantages)
    logp, ent = categorical(actor_output, torch.as_tensor(actions), is_probs)
    entropy = ent.mean()
    clipped = old_values + _tf_clip(values - old_values, -clip_norm, clip_norm)
    vl = 0.5 * _tf_max((values - returns) ** 2, (clipped - returns) ** 2).mean()
    ratio = torch.exp(logp - old_log_probs)
    pg = _tf_max(-advantages * ratio, -advantages * _tf_clip(ratio, 1 - clip_norm, 1 + clip_norm)).mean()
    loss = pg - entropy * entropy_coef + vl * value_loss_coef
    loss.backward()
    sc = dict(loss=loss.item(), pg=pg.item(), vl=vl.item(), entropy=entropy.item())
    return sc, actor_output.grad.numpy(), values.grad.numpy()


def a2c_loss_autograd(actor_output, values, actions, old_values, returns, entropy_coef, value_loss_coef,
                      is_probs=False, dtype=torch.float32):
    """xagents/a2c/agent.py:202-215 with autograd as the tape."""
    t = lambda x: torch.as_tensor(x, dtype=dtype)
    actor_output = t(actor_output).clone().requires_grad_(True)
    values = t(values).clone().requires_grad_(True)
    old_values, returns = t(old_values), t(returns)
    adv = returns - old_values
    logp, ent = categorical(actor_output, torch.as_tensor(actions), is_probs)
    entropy = ent.mean()
    pg = -(adv * logp).mean()
    vl = ((values - returns) ** 2).mean()
    loss = pg - entropy * entropy_coef + vl * value_loss_coef
    loss.backward()
    sc = dict(loss=loss.item(), pg=pg.item(), vl=vl.item(), entropy=entropy.item())
    return sc, actor_output.grad.numpy(), values.grad.numpy()


# ---- CPU-baseline timing legs (all host threads torch can use) ---------------------------------
def gather_rows(flat, idx):
    """tf.gather(item, batch_indices) (ppo/agent.py:154) on host memory."""
    return flat.index_select(0, idx)


def normalize_advantages(returns_mb, old_values_mb, eps):
    adv = returns_mb - old_values_mb
    return (adv - adv.mean()) / (adv.std(unbiased=False) + eps)


def ppo_loss_fwd_bwd(logits, values, actions, old_values, returns, old_log_probs, advantages,
                     clip_norm, entropy_coef, value_loss_coef):
    logits = logits.detach().requires_grad_(True)
    values = values.detach().requires_grad_(True)
    logp, ent = categorical(logits, actions)
    entropy = ent.mean()
    clipped = old_values + _tf_clip(values - old_values, -clip_norm, clip_norm)
    vl = 0.5 * _tf_max((values - returns) ** 2, (clipped - returns) ** 2).mean()
    ratio = torch.exp(logp - old_log_probs)
    pg = _tf_max(-advantages * ratio, -advantages * _tf_clip(ratio, 1 - clip_norm, 1 + clip_norm)).mean()
    loss = pg - entropy * entropy_coef + vl * value_loss_coef
    loss.backward()
    return loss.detach(), logits.grad, values.grad
