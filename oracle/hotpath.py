"""NumPy restatement of the reference's rollout-to-update arithmetic (checker only).

Every function cites the reference lines (relative to /root/reference) it follows and keeps
their floating-point operation ORDER, so an fp32 run of this file is what the reference's
NumPy / TF-CPU path computes.  `dtype=np.float64` gives the "truth twin" used for tolerance
accounting.  Shapes: T = n_steps, E = n_envs, N = T*E, flat index b = e*T + t (env-major).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import numpy as np

__all__ = [
    'gae_returns', 'nstep_returns', 'retrace_returns', 'concat_step_batches', 'env_major_to_time_major',
    'minibatch_slices', 'gather_minibatches', 'scale_images', 'normalize_advantages',
    'categorical_logp_entropy', 'diag_normal_logp_entropy', 'ppo_loss', 'ppo_loss_grads',
    'a2c_loss', 'a2c_loss_grads', 'clip_by_global_norm', 'adam_step', 'ppo_train_step',
]


# --------------------------------------------------------------------------- returns / GAE
def gae_returns(rewards, dones, values, next_values, gamma, lam, dtype=np.float32,
                return_advantages=False):
    """PPO returns = GAE advantages + values.  Follows xagents/ppo/agent.py:80-94.

    rewards [T,E]; dones [T+1,E] (row t+1 gates step t, a2c/agent.py:116,129,138);
    values [T,E]; next_values [E] stands in for the model call at ppo/agent.py:72-79.
    gamma*lam is folded as a Python float before it meets the array (ppo/agent.py:92).
    """
    rewards = np.asarray(rewards, dtype)
    dones = np.asarray(dones, dtype)
    values = np.asarray(values, dtype)
    boot = np.asarray(next_values, dtype).reshape(1, -1)
    n_steps = rewards.shape[0]
    all_values = np.concatenate([values, boot])                 # :82
    carry = 0
    gl = float(gamma) * float(lam)
    advantages = [None] * n_steps
    for t in range(n_steps - 1, -1, -1):                        # :84
        alive = 1 - dones[t + 1]                                # :85
        delta = rewards[t] + float(gamma) * all_values[t + 1] * alive - all_values[t]  # :87-91
        carry = delta + gl * alive * carry                      # :92
        advantages[t] = carry
    advantages = np.asarray(advantages)
    returns = advantages + all_values[:-1]                      # :94
    if return_advantages:
        return returns, advantages
    return returns


def nstep_returns(rewards, dones, next_values, gamma, dtype=np.float32):
    """A2C discounted n-step returns.  Follows xagents/a2c/agent.py:165-171."""
    rewards = np.asarray(rewards, dtype)
    dones = np.asarray(dones, dtype)
    running = np.asarray(next_values, dtype).reshape(-1)
    n_steps = rewards.shape[0]
    out = np.empty_like(rewards)
    one = dtype(1.0)
    for t in range(n_steps - 1, -1, -1):
        running = rewards[t] + (dtype(gamma) * running) * (one - dones[t + 1])   # :168-170
        out[t] = running
    return out


def retrace_returns(rewards, dones, values, last_values, q_selected, importance, gamma, dtype=np.float32):
    """ACER Retrace targets, time-major [T,E] in and out.  Follows xagents/acer/agent.py:198-208
    (dones[t] there is the flag AFTER step t, i.e. row t+1 of the [T+1,E] array used elsewhere here)."""
    rewards, values, q_selected = (np.asarray(x, dtype) for x in (rewards, values, q_selected))
    dones = np.asarray(dones, dtype)
    rho = np.minimum(dtype(1.0), np.asarray(importance, dtype))                 # :198
    current = np.asarray(last_values, dtype).reshape(-1)                        # values[-1], :203
    out = np.empty_like(rewards)
    for t in range(rewards.shape[0] - 1, -1, -1):
        current = rewards[t] + dtype(gamma) * current * (dtype(1.0) - dones[t + 1])   # :205
        out[t] = current
        current = (rho[t] * (current - q_selected[t])) + values[t]              # :207-209
    return out


# --------------------------------------------------------------------------- layout
def concat_step_batches(*fields):
    """Time-major [T,E,...] -> env-major flat [N,...].  Follows xagents/base.py:559-564."""
    flat = []
    for a in fields:
        a = np.asarray(a)
        if a.ndim == 1:
            a = a[:, None]
        flat.append(np.ascontiguousarray(np.swapaxes(a, 0, 1)).reshape((-1,) + a.shape[2:]))
    return flat


def env_major_to_time_major(b, n_steps, n_envs):
    """Row of the time-major [T*E] buffer that env-major flat index b = e*T + t names."""
    b = np.asarray(b, np.int64)
    return (b % n_steps) * n_envs + b // n_steps


def minibatch_slices(batch_size, mini_batch_size):
    """(start, stop) pairs of `range(0, N, B)` slicing incl. the trailing short one
    (ppo/agent.py:152-153)."""
    return [(i, min(i + mini_batch_size, batch_size)) for i in range(0, batch_size, mini_batch_size)]


def gather_minibatches(fields, permutations, mini_batch_size):
    """ppo/agent.py:149-154 with the shuffled index vectors supplied by the caller.

    fields: env-major flat arrays [N,...]; permutations: K int arrays of length N (the state of
    `indices` after each epoch's tf.random.shuffle).  Returns K*ceil(N/B) lists of gathered arrays.
    """
    out = []
    n = len(permutations[0])
    for perm in permutations:
        perm = np.asarray(perm)
        for lo, hi in minibatch_slices(n, mini_batch_size):
            out.append([np.take(f, perm[lo:hi], axis=0) for f in fields])   # tf.gather :154
    return out


def scale_images(states):
    """xagents/base.py:505-506: cast to fp32, true division by 255."""
    return np.asarray(states).astype(np.float32) / np.float32(255.0)


# --------------------------------------------------------------------------- advantages
def normalize_advantages(returns_mb, old_values_mb, eps, dtype=np.float32):
    """ppo/agent.py:180-183: adv = R - V_old; (adv - mean) / (population std + eps)."""
    adv = np.asarray(returns_mb, dtype) - np.asarray(old_values_mb, dtype)
    mean = adv.mean(dtype=dtype)
    std = np.sqrt(np.square(adv - mean).mean(dtype=dtype))     # tf.math.reduce_std, ddof 0
    return (adv - mean) / (std + dtype(eps))


# --------------------------------------------------------------------------- distributions
def _log_softmax(logits):
    z = logits - logits.max(axis=-1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def categorical_logp_entropy(actor_output, actions, is_probs=False, dtype=np.float32):
    """tfp Categorical(logits=) / (probs=) log_prob + entropy as used at a2c/agent.py:61-63,87,92.

    Published definition (tensorflow-probability 0.15.0, un-vendored): lsm = log_softmax(logits)
    (or log(probs)); log_prob(a) = lsm[a]; entropy = -sum(exp(lsm) * lsm).
    Returns (logp [n], entropy [n], lsm [n,A]).
    """
    x = np.asarray(actor_output, dtype)
    # probs= : tfp takes logits = log(probs) and still normalises through log_softmax
    with np.errstate(divide='ignore', invalid='ignore'):
        lsm = _log_softmax(np.log(x)) if is_probs else _log_softmax(x)
        a = np.asarray(actions).astype(np.int64).reshape(-1)
        logp = np.take_along_axis(lsm, a[:, None], axis=-1)[:, 0]
        p = np.exp(lsm)
        # tfp's Categorical.entropy uses multiply_no_nan: a probability of exactly 0 contributes 0 (not 0 * -inf)
        ent = -np.where(p > 0, p * lsm, dtype(0)).sum(axis=-1, dtype=dtype)
    return logp, ent, lsm


def diag_normal_logp_entropy(loc, actions, dtype=np.float32):
    """MultivariateNormalDiag(loc) with identity scale (a2c/agent.py:59-60)."""
    loc = np.asarray(loc, dtype)
    a = np.asarray(actions, dtype).reshape(loc.shape)
    k = loc.shape[-1]
    log2pi = dtype(np.log(2.0 * np.pi))
    logp = dtype(-0.5) * np.square(a - loc).sum(axis=-1, dtype=dtype) - dtype(0.5 * k) * log2pi
    ent = np.full(loc.shape[0], dtype(0.5 * k) * (dtype(1.0) + log2pi), dtype)
    return logp, ent


# --------------------------------------------------------------------------- losses
def ppo_loss(log_probs, values, entropy, old_values, returns, old_log_probs, advantages,
             clip_norm, entropy_coef, value_loss_coef, dtype=np.float32):
    """Forward of ppo/agent.py:116-133.  Returns dict(loss, pg, vl, entropy)."""
    f = lambda x: np.asarray(x, dtype)
    log_probs, values, entropy = f(log_probs), f(values), f(entropy)
    old_values, returns, old_log_probs, advantages = f(old_values), f(returns), f(old_log_probs), f(advantages)
    c = dtype(clip_norm)
    ent = entropy.mean(dtype=dtype)                                              # :116
    clipped_values = old_values + np.clip(values - old_values, -c, c)            # :117-119
    vl1 = np.square(values - returns)                                            # :120
    vl2 = np.square(clipped_values - returns)                                    # :121
    vl = dtype(0.5) * np.maximum(vl1, vl2).mean(dtype=dtype)                     # :122
    ratio = np.exp(log_probs - old_log_probs)                                    # :123
    pg1 = -advantages * ratio                                                    # :124
    pg2 = -advantages * np.clip(ratio, dtype(1) - c, dtype(1) + c)               # :125-127
    pg = np.maximum(pg1, pg2).mean(dtype=dtype)                                  # :128
    loss = pg - ent * dtype(entropy_coef) + vl * dtype(value_loss_coef)          # :129-133
    return dict(loss=loss, pg=pg, vl=vl, entropy=ent)


def ppo_loss_grads(actor_output, values, actions, old_values, returns, old_log_probs, advantages,
                   clip_norm, entropy_coef, value_loss_coef, is_probs=False, dtype=np.float32):
    """Closed-form d loss / d(actor_output, values) of ppo/agent.py:116-134 (what the tape
    hands to the model's backward).  TF tie rules: maximum -> first arg on ties; clip_by_value
    passes gradient on the closed interval."""
    f = lambda x: np.asarray(x, dtype)
    values, old_values, returns = f(values), f(old_values), f(returns)
    old_log_probs, advantages = f(old_log_probs), f(advantages)
    n = values.shape[0]
    c = dtype(clip_norm)
    logp, ent, lsm = categorical_logp_entropy(actor_output, actions, is_probs, dtype)
    p = np.exp(lsm)
    ratio = np.exp(logp - old_log_probs)
    s1 = -advantages * ratio
    s2 = -advantages * np.clip(ratio, dtype(1) - c, dtype(1) + c)
    g_logp = np.where(s1 >= s2, s1, dtype(0))            # d max / d logp (s1 = -adv*ratio, d ratio/d logp = ratio)
    # when s2 > s1 strictly the ratio sits outside the clip band -> clip gradient 0
    a = np.asarray(actions).astype(np.int64).reshape(-1)
    onehot = np.zeros_like(p)
    onehot[np.arange(n), a] = 1
    d_lsm = g_logp[:, None] * onehot + dtype(entropy_coef) * p * (lsm + dtype(1))   # dL/d lsm, *n
    d_actor = d_lsm - p * d_lsm.sum(axis=-1, keepdims=True, dtype=dtype)
    if is_probs:
        d_actor = d_actor / f(actor_output)              # chain through logits = log(probs)
    e1 = np.square(values - returns)
    vc = old_values + np.clip(values - old_values, -c, c)
    e2 = np.square(vc - returns)
    inside = np.abs(values - old_values) <= c
    d_e2 = np.where(inside, dtype(2) * (vc - returns), dtype(0))
    d_v = np.where(e1 >= e2, dtype(2) * (values - returns), d_e2) * dtype(0.5 * value_loss_coef)
    return d_actor / dtype(n), d_v / dtype(n)


def a2c_loss(log_probs, values, entropy, old_values, returns, entropy_coef, value_loss_coef,
             dtype=np.float32):
    """Forward of a2c/agent.py:202-214 (advantages un-normalised, :202)."""
    f = lambda x: np.asarray(x, dtype)
    log_probs, values, entropy, old_values, returns = map(f, (log_probs, values, entropy, old_values, returns))
    adv = returns - old_values                                                   # :202
    ent = entropy.mean(dtype=dtype)                                              # :207
    pg = -(adv * log_probs).mean(dtype=dtype)                                    # :208
    vl = np.square(values - returns).mean(dtype=dtype)                           # :209
    loss = pg - ent * dtype(entropy_coef) + vl * dtype(value_loss_coef)          # :210-214
    return dict(loss=loss, pg=pg, vl=vl, entropy=ent)


def a2c_loss_grads(actor_output, values, actions, old_values, returns, entropy_coef,
                   value_loss_coef, is_probs=False, dtype=np.float32):
    f = lambda x: np.asarray(x, dtype)
    values, old_values, returns = f(values), f(old_values), f(returns)
    n = values.shape[0]
    _, _, lsm = categorical_logp_entropy(actor_output, actions, is_probs, dtype)
    p = np.exp(lsm)
    adv = returns - old_values
    a = np.asarray(actions).astype(np.int64).reshape(-1)
    onehot = np.zeros_like(p)
    onehot[np.arange(n), a] = 1
    d_lsm = (-adv)[:, None] * onehot + dtype(entropy_coef) * p * (lsm + dtype(1))
    d_actor = d_lsm - p * d_lsm.sum(axis=-1, keepdims=True, dtype=dtype)
    if is_probs:
        d_actor = d_actor / f(actor_output)
    d_v = dtype(2 * value_loss_coef) * (values - returns)
    return d_actor / dtype(n), d_v / dtype(n)


# --------------------------------------------------------------------------- optimiser step
def clip_by_global_norm(grads, clip):
    """tf.clip_by_global_norm (ppo/agent.py:135-136): g * clip * min(1/norm, 1/clip)."""
    norm = np.sqrt(sum(np.square(g.astype(np.float32)).sum(dtype=np.float32) for g in grads)).astype(np.float32)
    scale = np.float32(clip) * np.minimum(np.float32(1) / norm, np.float32(1) / np.float32(clip))
    return [g * scale for g in grads], norm


def adam_step(param, grad, m, v, step, lr=7e-4, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras Adam (common.py:476, defaults utils/cli.py:14-25) as TF's dense ApplyAdam evaluates it:
    hyper-parameters are cast to fp32 first, so 1-beta is an fp32 subtraction; m += (g-m)(1-b1);
    v += (g^2-v)(1-b2); theta -= lr_t*m/(sqrt(v)+eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t)."""
    f = np.float32
    b1, b2 = f(beta1), f(beta2)
    m = m + (grad - m) * (f(1) - b1)
    v = v + (np.square(grad) - v) * (f(1) - b2)
    lr_t = f(lr) * np.sqrt(f(1) - b2 ** f(step)) / (f(1) - b1 ** f(step))
    return param - lr_t * m / (np.sqrt(v) + f(eps)), m, v


# --------------------------------------------------------------------------- whole path
def ppo_train_step(obs, rewards, dones, values, next_values, actions, old_log_probs,
                   permutations, new_logits, new_values, *, gamma=0.99, lam=0.95, mini_batches=4,
                   advantage_epsilon=1e-8, clip_norm=0.1, entropy_coef=0.01, value_loss_coef=0.5,
                   keep_states=False, dtype=np.float32):
    """One PPO.train_step over a pre-filled time-major rollout (ppo/agent.py:193-225) with the
    model replaced by supplied per-sample outputs `new_logits` [N,A] / `new_values` [N] indexed
    by env-major flat sample id (so that gathering them with the same minibatch indices plays the
    part of the forward pass at :113-115).

    Returns dict(returns [T,E], minibatches=[dict(idx, states?, loss, pg, vl, entropy, dlogits, dvalues)]).
    """
    n_steps, n_envs = np.asarray(rewards).shape
    returns = gae_returns(rewards, dones, values, next_values, gamma, lam, dtype)
    states, acts, rets, vals, logps = concat_step_batches(obs, actions, returns, values, old_log_probs)
    n = n_steps * n_envs
    mb_size = n // mini_batches
    out = []
    for perm in permutations:
        perm = np.asarray(perm)
        for lo, hi in minibatch_slices(n, mb_size):
            idx = perm[lo:hi]
            s_mb = np.take(states, idx, axis=0) if keep_states else None
            a_mb, r_mb, v_mb, lp_mb = (np.take(x, idx, axis=0).reshape(-1) for x in (acts, rets, vals, logps))
            adv = normalize_advantages(r_mb, v_mb, advantage_epsilon, dtype)
            logits_mb = np.take(new_logits, idx, axis=0)
            nv_mb = np.take(new_values, idx, axis=0)
            logp, ent, _ = categorical_logp_entropy(logits_mb, a_mb, False, dtype)
            scalars = ppo_loss(logp, nv_mb, ent, v_mb, r_mb, lp_mb, adv, clip_norm, entropy_coef,
                               value_loss_coef, dtype)
            dlogits, dvalues = ppo_loss_grads(logits_mb, nv_mb, a_mb, v_mb, r_mb, lp_mb, adv, clip_norm,
                                              entropy_coef, value_loss_coef, False, dtype)
            out.append(dict(idx=idx, states=s_mb, advantages=adv, dlogits=dlogits, dvalues=dvalues, **scalars))
    return dict(returns=returns, minibatches=out)
