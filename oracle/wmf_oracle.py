"""numpy restatement of Cornac's WMF training step (SURVEY.md section 8, row (f)4 -- groundwork for round 2).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED.  The reference implements WMF as a TensorFlow-1 graph (cornac/models/wmf/wmf.py:34-55 driven by
cornac/models/wmf/recom_wmf.py:152-212); TensorFlow is an un-vendored, unpinned dependency (wmf/requirements.txt is
empty) that is not installed here, so the reference cannot be run and there are no golden vectors: no reference test
asserts a WMF value either.  What follows restates the published semantics of the TF-1 ops the graph uses
(tf.nn.l2_loss = sum(x^2)/2, tf.clip_by_value, tf.train.AdamOptimizer with defaults beta1=0.9, beta2=0.999,
epsilon=1e-8 and its SPARSE update for the gathered variable V) and is checked only for internal consistency
(tests/test_oracle_golden.py: analytic gradients against finite differences of the restated loss, one Adam step
against a hand computation, monotone loss on a toy problem).

The graph (wmf.py:34-55), for one mini-batch of item ids `ids` (recom_wmf.py:186-199: R_b = R[:, ids] dense,
C_b = b everywhere, a where R_b != 0):
    V_b   = V[ids]                                   gather
    pred  = U @ V_b.T
    loss  = sum(C_b * (R_b - pred)^2) + lambda_u * sum(U^2)/2 + lambda_v * sum(V_b^2)/2
    grads = d loss / d U (dense), d loss / d V (IndexedSlices over `ids`)
    each gradient clipped elementwise to [-5, 5], then one Adam step on U and V.
"""
import numpy as np

F32 = np.float32


def xavier_uniform(shape, rng):
    """cornac/utils/init_utils.py:116-144: uniform(-l, l), l = sqrt(3) * sqrt(2 / (rows + cols)), f32."""
    std = np.sqrt(2.0 / np.sum(shape))
    limit = np.sqrt(3.0) * std
    return rng.uniform(-limit, limit, shape).astype(F32)


def batch_inputs(R_csc, ids, a, b):
    """recom_wmf.py:186-191: dense ratings of the batch's items and their confidence weights."""
    R_b = np.asarray(R_csc[:, ids].toarray(), dtype=F32)
    C_b = np.full(R_b.shape, b, dtype=F32)
    C_b[R_b != 0] = a
    return R_b, C_b


def loss_and_grads(U, V_b, R_b, C_b, lambda_u, lambda_v):
    """wmf.py:43-48 and the gradients TensorFlow derives from it (f32 arithmetic like the graph)."""
    pred = U @ V_b.T
    E = R_b - pred
    loss = F32(np.sum(C_b * E * E)) + F32(lambda_u) * F32(np.sum(U * U) / 2) + F32(lambda_v) * F32(np.sum(V_b * V_b) / 2)
    W = C_b * E                                        # [n_users, b]
    gU = F32(-2.0) * (W @ V_b) + F32(lambda_u) * U
    gVb = F32(-2.0) * (W.T @ U) + F32(lambda_v) * V_b
    return F32(loss), gU.astype(F32), gVb.astype(F32)


class AdamState:
    """Slots of tf.train.AdamOptimizer for one variable + the shared beta powers."""

    def __init__(self, shape):
        self.m = np.zeros(shape, F32)
        self.v = np.zeros(shape, F32)


class Adam:
    """tf.train.AdamOptimizer(lr) with TF-1 defaults.  `lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)`,
    `var -= lr_t * m / (sqrt(v) + epsilon)`; the beta powers advance once per apply_gradients call.
    apply_sparse follows _apply_sparse_shared: m and v of the WHOLE variable decay, the gradient rows are
    scatter-added, and the WHOLE variable moves (the non-lazy Adam of TF 1)."""

    def __init__(self, lr, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = F32(lr), F32(beta1), F32(beta2), F32(epsilon)
        self.b1_pow, self.b2_pow = F32(beta1), F32(beta2)

    def _lr_t(self):
        return self.lr * np.sqrt(F32(1) - self.b2_pow) / (F32(1) - self.b1_pow)

    def apply_dense(self, var, st, g):
        st.m[...] = self.b1 * st.m + (F32(1) - self.b1) * g
        st.v[...] = self.b2 * st.v + (F32(1) - self.b2) * g * g
        var -= self._lr_t() * st.m / (np.sqrt(st.v) + self.eps)

    def apply_sparse(self, var, st, ids, g_rows):
        st.m *= self.b1
        st.m[ids] += (F32(1) - self.b1) * g_rows
        st.v *= self.b2
        st.v[ids] += (F32(1) - self.b2) * g_rows * g_rows
        var -= self._lr_t() * st.m / (np.sqrt(st.v) + self.eps)

    def finish(self):
        self.b1_pow *= self.b1
        self.b2_pow *= self.b2


def train_step(U, V, ids, R_b, C_b, lambda_u, lambda_v, opt, st_U, st_V):
    """One sess.run([opt, loss]) of recom_wmf.py:197-199 (in place on U, V); returns the batch loss."""
    loss, gU, gVb = loss_and_grads(U, V[ids], R_b, C_b, lambda_u, lambda_v)
    gU = np.clip(gU, F32(-5), F32(5))                  # wmf.py:54
    gVb = np.clip(gVb, F32(-5), F32(5))
    opt.apply_dense(U, st_U, gU)
    opt.apply_sparse(V, st_V, np.asarray(ids), gVb)
    opt.finish()
    return loss


def fit(R_csc, U, V, batches, a=1.0, b=0.01, lambda_u=0.01, lambda_v=0.01, lr=0.001, max_iter=1):
    """WMF._fit_cf (recom_wmf.py:152-212) for an explicit batch schedule: `batches` is a callable returning the
    epoch's list of item-id arrays (train_set.item_iter(batch_size, shuffle=True) in the reference).
    Trains U, V in place; returns the mean batch loss per item of every epoch (the progress-bar figure)."""
    opt, st_U, st_V = Adam(lr), AdamState(U.shape), AdamState(V.shape)
    history = []
    for _ in range(int(max_iter)):
        total, count = 0.0, 0
        for ids in batches():
            R_b, C_b = batch_inputs(R_csc, ids, a, b)
            total += float(train_step(U, V, ids, R_b, C_b, lambda_u, lambda_v, opt, st_U, st_V))
            count += len(ids)
        history.append(total / max(count, 1))
    return history
