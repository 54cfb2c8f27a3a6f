"""The deferred-update scheme of `lazy_adam` as a numpy model, against the dense TF1 Adam of
the oracle, in float32 with the same operation order as adam_kernel / adam_rows_kernel.  This pins
the ALGORITHM on the CPU (the CUDA kernels are pinned against it on the GPU by tests/test_gpu_lazy_adam.py):
  * a row that is only brought up to date when a batch references it again -- one step with the gradient that
    was left in its gradient row, then zero-gradient steps -- ends with the same bits as a row updated densely;
  * the shortcut planned for long-idle rows is exact: once m is exactly 0 the parameter rests, once v is 0 too
    nothing changes any more.
"""
import numpy as np

F = np.float32


def lr_at(t, lr=1e-3, b1=0.9, b2=0.999):
    return F(float(lr) * np.sqrt(1.0 - float(b2) ** t) / (1.0 - float(b1) ** t))      # computed in double on the host, as the engine does


def dense_step(p, m, v, g, t, b1, b2, eps, lr=1e-3):
    """adam_kernel: every element, every step (correctly rounded fp32 operations in this order)."""
    b1, b2, eps = F(b1), F(b2), F(eps)
    m[:] = m * b1 + (F(1) - b1) * g
    v[:] = v * b2 + (F(1) - b2) * (g * g)
    p[:] = p - (lr_at(t, lr, float(b1), float(b2)) * m) / (np.sqrt(v) + eps)


class LazyTable:
    """Rows + slots + the bookkeeping of adam_rows_kernel: `last[r]` = step the row is current for, `g[r]` = the
    scatter-add of the step that last touched it (zeros otherwise)."""

    def __init__(self, p, b1, b2, eps, lr=1e-3, shortcut=False):
        self.p, self.m, self.v = p.copy(), np.zeros_like(p), np.zeros_like(p)
        self.g = np.zeros_like(p)
        self.last = np.zeros(p.shape[0], dtype=np.int64)
        self.t_done, self.hp, self.lr, self.shortcut = 0, (b1, b2, eps), lr, shortcut
        self.expensive_steps = 0

    def catch_up(self, rows):
        b1, b2, eps = self.hp
        for r in rows:
            if self.last[r] >= self.t_done:
                continue
            p, m, v = self.p[r:r + 1], self.m[r:r + 1], self.v[r:r + 1]
            s = self.last[r] + 1
            dense_step(p, m, v, self.g[r:r + 1], s, b1, b2, eps, self.lr)               # the deferred gradient step
            self.expensive_steps += 1
            self.g[r] = 0
            zero = np.zeros_like(p)
            s += 1
            while s <= self.t_done:
                if self.shortcut and not m.any():                                       # m == +-0 everywhere: theta rests
                    break
                dense_step(p, m, v, zero, s, b1, b2, eps, self.lr)
                self.expensive_steps += 1
                s += 1
            while s <= self.t_done:                                                     # only v still decays
                if not v.any():
                    break
                m[:] = m * F(b1) + (F(1) - F(b1)) * zero
                v[:] = v * F(b2) + (F(1) - F(b2)) * zero
                s += 1
            self.last[r] = self.t_done

    def train_step(self, rows, grads):
        self.catch_up(rows)                       # before the forward pass reads the rows
        np.add.at(self.g, rows, grads)            # backward: scatter-add
        self.t_done += 1                          # c2v_adam_step only records the step

    def flush(self):
        self.catch_up(range(self.p.shape[0]))


def _run(n_rows, d, steps, touch_prob, hp, seed, shortcut=False, lr=1e-3):
    rng = np.random.default_rng(seed)
    p0 = rng.standard_normal((n_rows, d)).astype(F)
    dense_p, dense_m, dense_v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    lazy = LazyTable(p0, *hp, lr=lr, shortcut=shortcut)
    for t in range(1, steps + 1):
        rows = np.flatnonzero(rng.random(n_rows) < touch_prob)
        grads = (rng.standard_normal((rows.size, d)) * 1e-2).astype(F)
        lazy.train_step(rows, grads)
        g = np.zeros_like(p0)
        g[rows] = grads
        dense_step(dense_p, dense_m, dense_v, g, t, *hp, lr=lr)
    lazy.flush()
    return lazy, dense_p, dense_m, dense_v


def test_deferred_updates_are_bit_identical_to_dense_adam():
    lazy, p, m, v = _run(n_rows=40, d=8, steps=60, touch_prob=0.25, hp=(0.9, 0.999, 1e-8), seed=0)
    assert np.array_equal(lazy.p, p) and np.array_equal(lazy.m, m) and np.array_equal(lazy.v, v)
    assert not lazy.g.any()                       # every gradient row was consumed and cleared


def test_long_idle_shortcut_is_exact_and_bounds_the_work():
    # small betas: m underflows to 0 after ~150 idle steps, v after ~900 -- most of 1200 steps are idle
    hp = (0.5, 0.9, 1e-8)
    with np.errstate(under="ignore"):
        plain, p, m, v = _run(n_rows=12, d=4, steps=1200, touch_prob=0.002, hp=hp, seed=3, lr=1e-2)
        quick, *_ = _run(n_rows=12, d=4, steps=1200, touch_prob=0.002, hp=hp, seed=3, shortcut=True, lr=1e-2)
    for a, b in ((plain.p, p), (plain.m, m), (plain.v, v), (quick.p, p), (quick.m, m), (quick.v, v)):
        assert np.array_equal(a, b)
    assert (np.signbit(quick.m) == np.signbit(m)).all()          # even the sign of a zero survives
    assert quick.expensive_steps < 0.4 * plain.expensive_steps   # the division / square-root loop was left early
