"""CPU statement of the deferred softmax normalisation (engine option exp_slab), checked against the
oracle's softmax cross entropy and its gradients: the algebra the CUDA kernels implement (umma::EpiExpSumT,
expsum_combine_kernel, expsum_rows_kernel / expsum_finish_kernel, scale_rows_kernel, the row factor in the dv reduction),
in numpy, single table and row-sharded.  The GPU tests compare the kernels with the two-pass schedule; this file pins
the scheme itself."""
import numpy as np
import pytest

from oracle import path_attention_oracle as O

EXP_SLAB_MIN, EXP_SLAB_MAX = np.float32(1e-26), np.float32(1e30)      # kExpSlabMin / kExpSlabMax in csrc/common.cuh


def two_pass(v, Y, target):
    """What the oracle (and the fallback schedule) computes: logits -> loss, G = (softmax - onehot) / B, dv = G Y, dY = G^T v."""
    B = v.shape[0]
    s = v @ Y.T
    _, per, lse = O.softmax_xent(s, target)
    G = np.exp(s - lse[:, None])
    G[np.arange(B), target] -= 1.0
    G /= B
    return per, lse, G @ Y, G.T @ v


def deferred(v, Y, target, dtype):
    """exp_slab on one table: c_b = true logit, U = exp(s - c), Z = sum U, one patched element and one factor per row."""
    B = v.shape[0]
    v, Y = v.astype(dtype), Y.astype(dtype)
    s = v @ Y.T
    c = np.einsum("bd,bd->b", v, Y[target])                  # true_logit_kernel
    U = np.exp(s - c[:, None])                                # EpiExpSumT
    Z = U.sum(axis=1)
    in_window = (U.max(axis=1) >= EXP_SLAB_MIN) & (U.max(axis=1) <= EXP_SLAB_MAX) & np.isfinite(Z)
    loss_b, lse = np.log(Z), c + np.log(Z)                    # expsum_combine_kernel
    U[np.arange(B), target] -= Z
    r = (dtype(1.0) / B) / Z
    dv = r[:, None] * (U @ Y)                                 # row factor applied by the split-K reduction
    dY = U.T @ (r[:, None] * v)                               # scale_rows_kernel feeds dY's small operand
    return loss_b, lse, dv, dY, in_window


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


@pytest.mark.parametrize("scale", [0.05, 1.0, 6.0])
def test_deferred_normalisation_is_the_softmax_gradient(scale):
    """float64: the two formulations agree to rounding, from a near-uniform softmax (scale 0.05: Z ~ |Y|) to a peaked one."""
    rng = np.random.default_rng(3)
    B, D, Yn = 37, 52, 1537
    v = rng.standard_normal((B, D))
    Y = rng.standard_normal((Yn, D)) * scale
    target = rng.integers(0, Yn, size=B)
    per, lse, dv, dY = two_pass(v, Y, target)
    loss_b, lse2, dv2, dY2, ok = deferred(v, Y, target, np.float64)
    assert ok.all() or scale > 1.0          # the window is fp32's; float64 carries the scale-6 rows anyway
    assert np.abs(loss_b - per).max() < 1e-10 and np.abs(lse2 - lse).max() < 1e-10
    assert rel(dv2, dv) < 1e-12 and rel(dY2, dY) < 1e-12


def test_deferred_normalisation_in_fp32_matches_fp32_two_pass():
    """fp32 storage and arithmetic: a row scaled by a constant rounds like the unscaled row, so the deferred form is as close
    to the float64 truth as the two-pass form is."""
    rng = np.random.default_rng(4)
    B, D, Yn = 64, 96, 4099
    v = rng.standard_normal((B, D)).astype(np.float32)
    Y = (rng.standard_normal((Yn, D)) * 0.3).astype(np.float32)
    target = rng.integers(0, Yn, size=B)
    per64, _, dv64, dY64 = two_pass(v.astype(np.float64), Y.astype(np.float64), target)
    per32, _, dv32, dY32 = two_pass(v, Y, target)
    loss_b, _, dv, dY, ok = deferred(v, Y, target, np.float32)
    assert ok.all() and dv.dtype == np.float32
    assert np.abs(loss_b - per64).max() < 2e-5
    assert rel(dv, dv64) < 4 * max(rel(dv32, dv64), 1e-6)
    assert rel(dY, dY64) < 4 * max(rel(dY32, dY64), 1e-6)


def test_range_guard_catches_rows_outside_the_fp32_window():
    """Logits hundreds of units above the true class: exp overflows, the guard (largest U of the row) must say so -- the
    engine then redoes the step with the two-pass kernels.  Rows inside the window are not flagged."""
    rng = np.random.default_rng(5)
    B, D, Yn = 16, 32, 512
    v = rng.standard_normal((B, D)).astype(np.float32)
    Y = rng.standard_normal((Yn, D)).astype(np.float32)
    target = rng.integers(0, Yn, size=B)
    with np.errstate(over="ignore", invalid="ignore"):
        *_, ok_calm = deferred(v, Y, target, np.float32)
        *_, ok_wild = deferred(v, Y * np.float32(40.0), target, np.float32)
    assert ok_calm.all()
    s = (v @ (Y * np.float32(40.0)).T).astype(np.float64)
    beyond = (s.max(axis=1) - s[np.arange(B), target]) > np.log(float(EXP_SLAB_MAX))
    assert beyond.any()
    assert not ok_wild[beyond].any()          # every row whose largest exponent exceeds the window is flagged


@pytest.mark.parametrize("world", [2, 8])
def test_row_sharded_combine_and_factors(world):
    """Row-sharded table: rank r holds classes [lo_r, hi_r), uses c_b^r = true logit if it owns y_b else 0, ships (c_b^r, Z_b^r)
    through the existing (max, sum) combine and, given the global log-sum-exp, patches its element by exp(lse_b - c_b^r)
    and scales by exp(c_b^r - lse_b) / B.  The summed dv partials and the concatenated dY blocks are the unsharded ones."""
    rng = np.random.default_rng(6)
    B, D, Yn = 24, 40, 1001
    v = rng.standard_normal((B, D))
    Y = rng.standard_normal((Yn, D)) * 0.7
    target = rng.integers(0, Yn, size=B)
    per, lse, dv, dY = two_pass(v, Y, target)
    bounds = [(r * Yn // world, (r + 1) * Yn // world) for r in range(world)]
    cs, Zs, Us = [], [], []
    for lo, hi in bounds:
        own = (target >= lo) & (target < hi)
        c = np.where(own, np.einsum("bd,bd->b", v, Y[np.clip(target, lo, hi - 1)]), 0.0)     # true_logit_kernel(row0, Y_local)
        U = np.exp(v @ Y[lo:hi].T - c[:, None])
        cs.append(c); Zs.append(U.sum(axis=1)); Us.append(U)
    cs, Zs = np.array(cs), np.array(Zs)
    m = cs.max(axis=0)                                                  # lse_combine_kernel on (c, Z) in place of (max, sum exp)
    lse2 = m + np.log((Zs * np.exp(cs - m)).sum(axis=0))
    assert np.abs(lse2 - lse).max() < 1e-10
    true_logit = np.einsum("bd,bd->b", v, Y[target])                    # all-reduce of the owners' values
    assert np.abs((lse2 - true_logit) - per).max() < 1e-10
    dv2, dY2 = np.zeros_like(dv), np.zeros_like(dY)
    for (lo, hi), c, U in zip(bounds, cs, Us):
        own = (target >= lo) & (target < hi)
        d = c - lse2
        assert (np.abs(d) < 80).all()
        U = U.copy()
        U[np.flatnonzero(own), target[own] - lo] -= np.exp(-d[own])    # expsum_finish_kernel
        r = np.exp(d) / B
        dv2 += r[:, None] * (U @ Y[lo:hi])                              # reduce-scatter sums the ranks' partials
        dY2[lo:hi] = U.T @ (r[:, None] * v)
    assert rel(dv2, dv) < 1e-12 and rel(dY2, dY) < 1e-12
