"""TEST INFRASTRUCTURE -- CPU restatement of the brains' arithmetic (never imported by reinlife_b200/).

The reference's network math lives in an un-vendored third-party dependency: PyTorch
(`torch>=1.3.1`, requirements.txt:1; this container: torch 2.11.0 CPU).  This file restates, with
explicit fp32 tensor algebra and NO autograd / nn.Module / optimizer objects:

* the three forwards            Models/PERD3QN.py:198-202, Models/DQN.py:126-130, Models/PPO.py:101-112
* one PERD3QN / D3QN train()    Models/PERD3QN.py:94-115, Models/D3QN.py:97-116 (MSE on r + g(1-d)max_a Q_target,
                                priorities |max_a Q_target - Q(s,a)|, whole-tensor advantage mean)
* torch.optim.Adam defaults     (betas .9/.999, eps 1e-8, bias-corrected)
* the action rules              Models/PERD3QN.py:204-210, Models/DQN.py:132-139
* the event-batched update used for N worlds (mean of per-event gradients, one Adam step)
* the integer-CDF proportional sampler that stands in for np.random.choice(len, 64, p) (PERD3QN.py:157-165)

Pinned against tests/golden/brain_golden.npz, which oracle/make_brain_golden.py mints by running the
reference's own modules (DuelingDDQN / Qnet / PPO / PERD3QNAgent.train with the pretrained weights).
State dicts use the reference's key names.
"""
import numpy as np
import torch

F32 = torch.float32


def _t(x):
    return torch.as_tensor(np.asarray(x), dtype=F32)


def lin(x, w, b):
    return x @ _t(w).T + _t(b)


# ------------------------------------------------------------------ forwards
def dueling_parts(sd, x):
    x = _t(x)
    feat = lin(x, sd["fc.weight"], sd["fc.bias"])
    h1 = torch.relu(feat)
    a1 = torch.relu(lin(h1, sd["adv_fc1.weight"], sd["adv_fc1.bias"]))
    v1 = torch.relu(lin(h1, sd["value_fc1.weight"], sd["value_fc1.bias"]))
    adv = lin(a1, sd["adv_fc2.weight"], sd["adv_fc2.bias"])
    val = lin(v1, sd["value_fc2.weight"], sd["value_fc2.bias"])
    return x, h1, a1, v1, adv, val


def dueling_forward(sd, x, per_row_mean):
    """per_row_mean=True: B=1 act semantics (each row its own batch); False: train semantics (PERD3QN.py:202)."""
    *_, adv, val = dueling_parts(sd, x)
    mean = adv.mean(1, keepdim=True) if per_row_mean else adv.mean()
    return (adv + val - mean).numpy()


def dqn_forward(sd, x):
    x = _t(x)
    h = torch.relu(lin(x, sd["fc1.weight"], sd["fc1.bias"]))
    h = torch.relu(lin(h, sd["fc2.weight"], sd["fc2.bias"]))
    return lin(h, sd["fc3.weight"], sd["fc3.bias"]).numpy()


def ppo_forward(sd, x):
    x = _t(x)
    h = torch.relu(lin(x, sd["fc1.weight"], sd["fc1.bias"]))
    h = torch.relu(lin(h, sd["fc2.weight"], sd["fc2.bias"]))
    logits = lin(h, sd["fc_pi.weight"], sd["fc_pi.bias"])
    z = logits - logits.max(1, keepdim=True)[0]
    e = torch.exp(z)
    return (e / e.sum(1, keepdim=True)).numpy(), lin(h, sd["fc_v.weight"], sd["fc_v.bias"]).numpy()


# ------------------------------------------------------------------ one dueling train() event, explicit backward
def dueling_event_grads(sd_eval, sd_target, obs, action, reward, next_obs, done, gamma):
    """Returns (grads dict in state_dict orientation, loss, priorities) for ONE 64-row event."""
    B = len(action)
    x, h1, a1, v1, adv, val = dueling_parts(sd_eval, obs)
    q = adv + val - adv.mean()
    qn = torch.as_tensor(dueling_forward(sd_target, next_obs, per_row_mean=False))
    next_q = qn.max(1)[0]
    a = torch.as_tensor(np.asarray(action), dtype=torch.long)
    q_a = q.gather(1, a[:, None])[:, 0]
    y = _t(reward) + gamma * (1 - _t(done)) * next_q
    loss = ((q_a - y) ** 2).mean()
    prio = (next_q - q_a).abs()
    g = 2 * (q_a - y) / B                                   # dL/dQ[b, a_b]
    d_adv = torch.zeros(B, 8)
    d_adv[torch.arange(B), a] = g
    d_adv -= g.sum() / (8 * B)                              # whole-tensor mean couples every row (Appendix C)
    d_val = g[:, None]
    grads = {}
    grads["adv_fc2.weight"], grads["adv_fc2.bias"] = d_adv.T @ a1, d_adv.sum(0)
    grads["value_fc2.weight"], grads["value_fc2.bias"] = d_val.T @ v1, d_val.sum(0)
    d_a1 = (d_adv @ _t(sd_eval["adv_fc2.weight"])) * (a1 > 0)
    d_v1 = (d_val @ _t(sd_eval["value_fc2.weight"])) * (v1 > 0)
    grads["adv_fc1.weight"], grads["adv_fc1.bias"] = d_a1.T @ h1, d_a1.sum(0)
    grads["value_fc1.weight"], grads["value_fc1.bias"] = d_v1.T @ h1, d_v1.sum(0)
    d_h1 = (d_a1 @ _t(sd_eval["adv_fc1.weight"]) + d_v1 @ _t(sd_eval["value_fc1.weight"])) * (h1 > 0)
    grads["fc.weight"], grads["fc.bias"] = d_h1.T @ x, d_h1.sum(0)
    return {k: v.numpy() for k, v in grads.items()}, float(loss), prio.numpy()


def adam_step(params, grads, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (defaults, no amsgrad / weight decay), in place on dicts of float32 numpy arrays. `step` is 1-based."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    for k in params:
        g = _t(grads[k])
        mk = _t(m[k]); vk = _t(v[k]); p = _t(params[k])
        mk = mk + (g - mk) * (1 - beta1)                    # exp_avg.lerp_(grad, 1-beta1)
        vk = vk * beta2 + (1 - beta2) * g * g               # exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1-beta2)
        denom = vk.sqrt() / np.sqrt(bc2) + eps
        p = p - step_size * (mk / denom)
        m[k], v[k], params[k] = mk.numpy(), vk.numpy(), p.numpy()


def dueling_batched_update(sd_eval, sd_target, events, gamma):
    """N-world semantics: mean over events of the per-event gradient (DESIGN.md 'learn step')."""
    acc, losses, prios = None, [], []
    for ev in events:
        g, loss, prio = dueling_event_grads(sd_eval, sd_target, *ev, gamma)
        losses.append(loss); prios.append(prio)
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    return {k: acc[k] / len(events) for k in acc}, losses, prios


# ------------------------------------------------------------------ action rules (given network outputs)
def first_argmax(q):
    return int(np.argmax(np.asarray(q)))   # first maximum, like torch.max / Tensor.argmax on CPU


def dueling_rule(q, eps, u_explore, r_below8):
    return first_argmax(q) if u_explore > eps else int(r_below8)      # PERD3QN.py:204-210


def dqn_rule(q, eps, coin, r_below8):
    return int(r_below8) if coin < eps else first_argmax(q)           # DQN.py:135-139


# ------------------------------------------------------------------ proportional sampler (integer CDF)
def per_weight(prio):
    """float32(float64(p) ** 0.6): the alpha-power of PERD3QN.py:162, rounded once."""
    return np.power(np.asarray(prio, np.float64), 0.6).astype(np.float32)


def per_sample(weights_f32, u53):
    """weights: float32 [len]; u53: python ints < 2**53 (top 53 bits of each draw).  Exact integer arithmetic:
    fixed-point weights w*2^24, inclusive prefix sums, index = first i with cum[i] > floor(u53 * total / 2^53)."""
    fix = [int(np.float64(w) * 16777216.0) for w in weights_f32]
    cum, run = [], 0
    for f in fix:
        run += f
        cum.append(run)
    total = run
    out = []
    for u in u53:
        if total == 0:
            out.append(0)
            continue
        target = (int(u) * total) >> 53
        lo, hi = 0, len(cum) - 1
        while lo < hi:
            mid = (lo + hi) // 2
            if cum[mid] > target:
                hi = mid
            else:
                lo = mid + 1
        out.append(lo)
    return out


# ------------------------------------------------------------------ uniform sampler without replacement
def uniform_sample(n, k, bits_fn):
    """random.sample(population, k) of CPython (Lib/random.py, Random.sample; used by D3QN.py:140 and DQN.py:100),
    restated on a counter generator: the c-th _randbelow(m) call returns (bits_fn(c) >> 32) * m >> 32.
    Returns k distinct population indices (0 = oldest item of the deque).  n >= k is the caller's business
    (the reference raises ValueError otherwise)."""
    import math
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")
    c = 0

    def randbelow(m):
        nonlocal c
        v = ((bits_fn(c) >> 32) * m) >> 32
        c += 1
        return v
    setsize = 21
    if k > 5:
        setsize += 4 ** math.ceil(math.log(k * 3, 4))
    result = [None] * k
    if n <= setsize:                       # pool method
        pool = list(range(n))
        for i in range(k):
            j = randbelow(n - i)
            result[i] = pool[j]
            pool[j] = pool[n - i - 1]
    else:                                  # set method with redraws
        selected = set()
        for i in range(k):
            j = randbelow(n)
            while j in selected:
                j = randbelow(n)
            selected.add(j)
            result[i] = j
    return result


# ------------------------------------------------------------------ DQN: one iteration of train() (Models/DQN.py:142-153)
def _mlp3_parts(sd, x, names):
    x = _t(x)
    h1 = torch.relu(lin(x, sd[names[0] + ".weight"], sd[names[0] + ".bias"]))
    h2 = torch.relu(lin(h1, sd[names[1] + ".weight"], sd[names[1] + ".bias"]))
    return x, h1, h2


def _mlp3_backward(sd, names, x, h1, h2, d_h2):
    """Given dL/dh2 (post-ReLU activations), the gradients of the two trunk layers."""
    g = {}
    d_z2 = d_h2 * (h2 > 0)
    g[names[1] + ".weight"], g[names[1] + ".bias"] = d_z2.T @ h1, d_z2.sum(0)
    d_z1 = (d_z2 @ _t(sd[names[1] + ".weight"])) * (h1 > 0)
    g[names[0] + ".weight"], g[names[0] + ".bias"] = d_z1.T @ x, d_z1.sum(0)
    return g


def dqn_iter_grads(sd, sd_target, obs, action, reward, next_obs, done_mask, gamma=0.98):
    """(grads, loss) of ONE of the five iterations: smooth_l1_loss(q(s)[a], r + gamma * max q_target(s') * done_mask),
    mean over the batch, beta = 1."""
    B = len(action)
    x, h1, h2 = _mlp3_parts(sd, obs, ("fc1", "fc2"))
    q = lin(h2, sd["fc3.weight"], sd["fc3.bias"])
    a = torch.as_tensor(np.asarray(action), dtype=torch.long)
    q_a = q.gather(1, a[:, None])[:, 0]
    max_q = torch.as_tensor(dqn_forward(sd_target, next_obs)).max(1)[0]
    y = _t(reward) + gamma * max_q * _t(done_mask)
    d = q_a - y
    loss = torch.where(d.abs() < 1, 0.5 * d * d, d.abs() - 0.5).mean()
    g_q = d.clamp(-1, 1) / B
    d_out = torch.zeros(B, 8)
    d_out[torch.arange(B), a] = g_q
    grads = {"fc3.weight": d_out.T @ h2, "fc3.bias": d_out.sum(0)}
    grads.update(_mlp3_backward(sd, ("fc1", "fc2"), x, h1, h2, d_out @ _t(sd["fc3.weight"])))
    return {k: v.numpy() for k, v in grads.items()}, float(loss)


# ------------------------------------------------------------------ PPO: one epoch of learn() (Models/PPO.py:136-162)
def ppo_gae(delta, gamma, lmbda):
    """PPO.py:143-150 under numpy >= 2 (NEP 50): `gamma * lmbda` is a python float, `advantage` becomes np.float32
    after the first addition, so every step is float32(float32(gamma*lmbda) * adv) + delta_t in float32."""
    gl = np.float32(gamma * lmbda)
    adv = np.zeros(len(delta), np.float32)
    run = np.float32(0.0)
    for t in range(len(delta) - 1, -1, -1):
        run = np.float32(np.float32(gl * run) + np.float32(delta[t]))
        adv[t] = run
    return adv


def ppo_epoch_grads(sd, obs, action, reward, next_obs, prob_a, done, gamma=0.98, lmbda=0.95, eps_clip=0.1):
    """(grads, loss) of ONE epoch on ONE data list of T transitions: loss.mean() with
    loss = -min(ratio*A, clamp(ratio, 1-eps, 1+eps)*A) + smooth_l1_loss(v(s), td_target)  (the second term a scalar mean)."""
    T = len(action)
    names = ("fc1", "fc2")
    x, h1, h2 = _mlp3_parts(sd, obs, names)
    logits = lin(h2, sd["fc_pi.weight"], sd["fc_pi.bias"])
    v = lin(h2, sd["fc_v.weight"], sd["fc_v.bias"])[:, 0]
    _, _, h2n = _mlp3_parts(sd, next_obs, names)
    v_next = lin(h2n, sd["fc_v.weight"], sd["fc_v.bias"])[:, 0]
    done_mask = _t(1.0 - np.asarray(done, np.float32))
    td = _t(reward) + gamma * v_next * done_mask
    adv = _t(ppo_gae((td - v).numpy(), gamma, lmbda))
    z = logits - logits.max(1, keepdim=True)[0]
    pi = torch.exp(z) / torch.exp(z).sum(1, keepdim=True)
    a = torch.as_tensor(np.asarray(action), dtype=torch.long)
    pi_a = pi.gather(1, a[:, None])[:, 0]
    ratio = torch.exp(torch.log(pi_a) - torch.log(_t(prob_a)))
    lo, hi = np.float32(1 - eps_clip), np.float32(1 + eps_clip)
    surr1, surr2 = ratio * adv, ratio.clamp(lo, hi) * adv
    dv = v - td
    sl1 = torch.where(dv.abs() < 1, 0.5 * dv * dv, dv.abs() - 0.5).mean()
    loss = (-torch.min(surr1, surr2)).mean() + sl1
    inside = (ratio >= lo) & (ratio <= hi)
    d_ratio = -(adv / T) * (inside | (surr1 < surr2))
    d_logpa = d_ratio * ratio
    onehot = torch.zeros(T, 8)
    onehot[torch.arange(T), a] = 1
    d_logits = d_logpa[:, None] * (onehot - pi)
    d_v = dv.clamp(-1, 1) / T
    grads = {"fc_pi.weight": d_logits.T @ h2, "fc_pi.bias": d_logits.sum(0),
             "fc_v.weight": d_v[None, :] @ h2, "fc_v.bias": d_v.sum(0, keepdim=True)}
    d_h2 = d_logits @ _t(sd["fc_pi.weight"]) + d_v[:, None] @ _t(sd["fc_v.weight"])
    grads.update(_mlp3_backward(sd, names, x, h1, h2, d_h2))
    return {k: v_.numpy() for k, v_ in grads.items()}, float(loss)


# ------------------------------------------------------------------ PERDQN (Models/PERDQN.py)
PERDQN_NAMES = ("fc.0", "fc.2", "fc.4")          # nn.Sequential indices of the three Linear layers (PERDQN.py:314-320)


def perdqn_forward(sd, x):
    """DQN.forward, PERDQN.py:311-323: 153 -> 64 -> 64 -> 8, ReLU between."""
    _, _, h2 = _mlp3_parts(sd, x, PERDQN_NAMES[:2])
    return lin(h2, sd["fc.4.weight"], sd["fc.4.bias"]).numpy()


def perdqn_rule(q, eps, u, r_below8):
    """get_action, PERDQN.py:101-111: np.random.rand() <= eps -> random.randrange(8), else torch.max(q, 1) index."""
    return int(r_below8) if u <= eps else first_argmax(q)


def perdqn_store_error(sd, sd_target, state, action, reward, next_state, done, gamma=0.99):
    """append_sample, PERDQN.py:113-128.  `old_val = target[0][action]` is a VIEW of the tensor that the next lines
    overwrite in place (`target[0][action] = reward [+ gamma * max Q_target(s')]`), so
    `error = abs(old_val - target[0][action])` compares the new value with itself: the stored error is exactly 0 for every
    transition, whatever the two B=1 forwards computed (pinned by the golden run: all 400 recorded errors are 0.0).
    New items therefore always enter the tree with priority float32(0.01) ** 0.6."""
    return np.zeros(len(np.asarray(action)), np.float32)


def perdqn_priority(error_f32, e=0.01, a=0.6):
    """Memory._get_priority, PERDQN.py:272-273.  The error reaches it as a float32 torch scalar (add, :126-128) or a
    numpy float32 scalar (update, :168-174); with numpy >= 2 (NEP 50) `(abs(error) + 0.01) ** 0.6` stays float32 in both
    cases (numpy 1.x promoted the numpy-scalar case to float64).  The tree then holds float64(float32 value)."""
    # scalar by scalar, as the reference does: numpy's float32 SCALAR power is libm powf, its array power a SIMD routine
    # that differs in the last bit for ~15% of the inputs
    flat = np.asarray(error_f32, np.float32).reshape(-1)
    out = np.array([(np.abs(x) + e) ** a for x in flat], np.float32)
    return out.reshape(np.shape(error_f32))


class SumTreeOracle:
    """SumTree + Memory of PERDQN.py:198-308 without the python objects: `tree` float64 [2*capacity-1] with the
    reference's incremental `+= change` propagation (so rounding matches), ring write pointer, n_entries, beta."""

    def __init__(self, capacity, beta=0.4, beta_increment=0.001):
        self.capacity = int(capacity)
        self.tree = np.zeros(2 * self.capacity - 1)
        self.write, self.n_entries = 0, 0
        self.beta, self.beta_increment = beta, beta_increment

    def update(self, idx, p):                      # SumTree.update + _propagate, :243-247, :211-218 (Memory.update path)
        """`p` arrives as a numpy float32 scalar (train_model, :168-174): numpy-scalar arithmetic, float64 throughout."""
        change = np.float64(p) - self.tree[idx]
        self.tree[idx] = p
        while idx != 0:
            idx = (idx - 1) // 2
            self.tree[idx] += change

    def _update_f32(self, idx, p):
        """The same two functions when `p` is a 0-d float32 torch tensor (Memory.add <- append_sample, :126-128):
        `p - self.tree[idx]` and `self.tree[parent] += change` are then TENSOR operations (numpy defers to
        Tensor.__rsub__/__radd__), i.e. float32 arithmetic on the float32-rounded node, stored back as float64."""
        f = np.float32
        change = f(f(p) - f(self.tree[idx]))
        self.tree[idx] = np.float64(f(p))
        while idx != 0:
            idx = (idx - 1) // 2
            self.tree[idx] = np.float64(f(f(self.tree[idx]) + change))

    def add(self, p):                              # SumTree.add, :229-240; returns the data slot written
        slot = self.write
        self._update_f32(slot + self.capacity - 1, p)
        self.write = (self.write + 1) % self.capacity
        self.n_entries = min(self.capacity, self.n_entries + 1)
        return slot

    def retrieve(self, s):                         # SumTree._retrieve, :220-230, iterative
        idx = 0
        while True:
            left = 2 * idx + 1
            if left >= len(self.tree):
                return idx
            if s <= self.tree[left]:
                idx = left
            else:
                s = s - self.tree[left]
                idx = left + 1

    def sample(self, n, next_u=None, max_tries=1 << 20, indexed_u=None):
        """Memory.sample, :278-303.  `next_u()` returns the float of the next random.random() call
        (random.uniform(a, b) = a + (b - a) * random()); alternatively `indexed_u(stratum, redraw)` serves the draws of
        a counter generator.  Returns (data slots, tree indices, is_weights float64)."""
        total = self.tree[0]
        segment = total / n
        self.beta = float(np.min([1., self.beta + self.beta_increment]))
        slots, idxs, prios = [], [], []
        for i in range(n):
            a, b = segment * i, segment * (i + 1)
            for tries in range(max_tries):
                s = a + (b - a) * (indexed_u(i, tries) if indexed_u is not None else next_u())
                idx = self.retrieve(s)
                slot = idx - self.capacity + 1
                if slot < self.n_entries:          # `not isinstance(data, int)`: the slot has been written
                    break
            else:
                raise RuntimeError("stratum without a filled leaf")
            slots.append(slot); idxs.append(idx); prios.append(self.tree[idx])
        probs = np.asarray(prios) / total
        w = np.power(self.n_entries * probs, -self.beta)
        w /= w.max()
        return slots, idxs, w


def perdqn_event_grads(sd, sd_target, obs, action, reward, next_obs, done, is_weights, gamma=0.99):
    """train_model, PERDQN.py:130-186, given the sampled batch: (grads, loss, errors).  `F.mse_loss(pred, target)` is a
    scalar mean there, so loss = mean_i(float32(is_w_i) * mse) and every row's gradient carries mean(float32(is_w))."""
    B = len(action)
    x, h1, h2 = _mlp3_parts(sd, obs, PERDQN_NAMES[:2])
    q = lin(h2, sd["fc.4.weight"], sd["fc.4.bias"])
    a = torch.as_tensor(np.asarray(action), dtype=torch.long)
    pred = q.gather(1, a[:, None])[:, 0]
    nmax = torch.as_tensor(perdqn_forward(sd_target, next_obs)).max(1)[0]
    tgt = _t(reward) + (1 - _t(np.asarray(done, np.float32))) * gamma * nmax
    errors = (pred - tgt).abs()
    w = _t(np.asarray(is_weights, np.float32))
    mse = ((pred - tgt) ** 2).mean()
    loss = (w * mse).mean()
    g_q = (w.sum() / B) * 2 * (pred - tgt) / B
    d_out = torch.zeros(B, 8)
    d_out[torch.arange(B), a] = g_q
    grads = {"fc.4.weight": d_out.T @ h2, "fc.4.bias": d_out.sum(0)}
    grads.update(_mlp3_backward(sd, PERDQN_NAMES[:2], x, h1, h2, d_out @ _t(sd["fc.4.weight"])))
    return {k: v.numpy() for k, v in grads.items()}, float(loss), errors.numpy()
