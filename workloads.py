"""Graph recipes of the BASELINE.json configurations, built with the reference's
`age` graph API (shared by bench.py and the tests).

  chain   out = add(mul(exp(x), extend(b, 1, {D})), permute(x, {1,0}))     config 2
  reduce  reduce_sum / reduce_max along each axis of [D, D]                 config 3
  matmul  age.matmul(a, b)                                                  config 4
  mlp     784-1024-1024-10 sigmoid MLP, squared-error loss / batch,
          gradients of all six parameters by llo.derive                     config 5
"""
import numpy as np

MLP_SIZES = (784, 1024, 1024, 10)


def chain_inputs(side, dtype=np.float32, seed_x=2, seed_b=3):
    x = np.random.default_rng(seed_x).uniform(0.5, 2.0, (side, side)).astype(dtype)
    b = np.random.default_rng(seed_b).uniform(0.5, 2.0, (side,)).astype(dtype)
    return x, b


def chain_graph(age, llo, x, b):
    side = x.shape[0]
    vx = llo.variable(x, "x")
    vb = llo.variable(b, "b")
    root = age.add(age.mul(age.exp(vx), age.extend(vb, 1, [side])), age.permute(vx, [1, 0]))
    return vx, vb, root


def chain_bytes(side, itemsize=4):
    n = side * side
    return itemsize * (n + side + n)


def mlp_data(batch, sizes=MLP_SIZES, dtype=np.float32, seed=10):
    """numpy-order tensors of SURVEY.md 8d config 5 (seeds 10..17)"""
    rngs = [np.random.default_rng(seed + i) for i in range(2 + 2 * (len(sizes) - 1))]
    X = rngs[0].uniform(0.05, 1.0, (batch, sizes[0])).astype(dtype)
    Y = rngs[1].uniform(0.0, 1.0, (batch, sizes[-1])).astype(dtype)
    params = []
    for i in range(len(sizes) - 1):
        W = (rngs[2 + 2 * i].standard_normal((sizes[i], sizes[i + 1])) / np.sqrt(sizes[i])).astype(dtype)
        b = (0.1 * rngs[3 + 2 * i].standard_normal((sizes[i + 1],))).astype(dtype)
        params += [W, b]
    return X, Y, params


def mlp_graph(age, llo, X, Y, params, global_batch=None):
    """forward + loss over the rows of X / Y given (a shard or the whole
    batch); the loss is scaled by `global_batch` so shard gradients add up to
    the full-batch gradient.  sigmoid(z) = 1 / (1 + exp(-z)) is composed from
    the op table (no activation opcode exists in llo/cfg/llo.json)."""
    rows = X.shape[0]
    if global_batch is None:
        global_batch = rows
    vx = llo.variable(X, "X")
    vy = llo.variable(Y, "Y")
    pvars = []
    h = vx
    for i in range(len(params) // 2):
        W, b = params[2 * i], params[2 * i + 1]
        vw = llo.variable(W, "W%d" % (i + 1))
        vb = llo.variable(b, "b%d" % (i + 1))
        pvars += [vw, vb]
        z = age.add(age.matmul(h, vw), age.extend(vb, 1, [rows]))
        one = llo.scalar(1.0, [rows, W.shape[1]])
        h = age.div(one, age.add(one, age.exp(age.neg(z))))
    diff = age.sub(h, vy)
    sq = age.pow(diff, llo.scalar(2.0, [rows, Y.shape[1]]))
    loss = age.div(age.reduce_sum0(sq), llo.scalar(float(global_batch), []))
    return loss, pvars, (vx, vy)


def mlp_flops(batch, sizes=MLP_SIZES):
    """GEMM flops of one gradient evaluation: forward + dW for
    every layer, dX for every layer but the first"""
    f = [sizes[i] * sizes[i + 1] for i in range(len(sizes) - 1)]
    return 2.0 * batch * (2 * sum(f) + sum(f[1:]))


def mlp_reference_gradients(X, Y, params, global_batch=None):
    """float64 numpy gradients of the same loss (independent check)"""
    X = X.astype(np.float64)
    Y = Y.astype(np.float64)
    P = [p.astype(np.float64) for p in params]
    B = X.shape[0] if global_batch is None else global_batch
    hs = [X]
    for i in range(len(P) // 2):
        z = hs[-1] @ P[2 * i] + P[2 * i + 1]
        hs.append(1.0 / (1.0 + np.exp(-z)))
    loss = ((hs[-1] - Y) ** 2).sum() / B
    grads = [None] * len(P)
    delta = 2.0 * (hs[-1] - Y) / B
    for i in reversed(range(len(P) // 2)):
        dz = delta * hs[i + 1] * (1.0 - hs[i + 1])
        grads[2 * i] = hs[i].T @ dz
        grads[2 * i + 1] = dz.sum(axis=0)
        delta = dz @ P[2 * i].T
    return loss, grads
