"""Seeded synthetic TOPP-RA problems shared by tests, bench.py and smoke().

Generation order per path follows the reference example examples/plot_kinematics.py:22-33:
way_pts = randn(n, dof); vlims = 10 + rand(dof)*20; alims = 10 + rand(dof)*2; ss = linspace(0,1,n)."""
import numpy as np


def make_path(seed, dof=7, nway=5, vel_active=False):
    rng = np.random.RandomState(seed)
    way = rng.randn(nway, dof)
    vl = (0.5 + rng.rand(dof)) if vel_active else (10 + rng.rand(dof) * 20)
    al = 10 + rng.rand(dof) * 2
    vlim = np.vstack((-vl, vl)).T
    alim = np.vstack((-al, al)).T
    return way, vlim, alim


def make_batch(B, seed0=1000, dof=7, nway=5, vel_active=False):
    """B paths with seeds seed0+b.  Returns ss [n], way [B,n,dof], vlim [B,dof,2], alim [B,dof,2]."""
    way = np.empty((B, nway, dof))
    vlim = np.empty((B, dof, 2))
    alim = np.empty((B, dof, 2))
    for b in range(B):
        way[b], vlim[b], alim[b] = make_path(seed0 + b, dof, nway, vel_active)
    return np.linspace(0, 1, nway), way, vlim, alim


def make_batch_fast(B, seed=1234, dof=7, nway=5):
    """Large batches for the bench: one RandomState for the whole batch (same distributions)."""
    rng = np.random.RandomState(seed)
    way = rng.randn(B, nway, dof)
    vl = 10 + rng.rand(B, dof) * 20
    al = 10 + rng.rand(B, dof) * 2
    vlim = np.stack((-vl, vl), axis=-1)
    alim = np.stack((-al, al), axis=-1)
    return np.linspace(0, 1, nway), way, vlim, alim


# ---- cfg 3: synthetic closed-form torque model:
#      tau = M(q) qdd + h(q) |qd|^2 + g(q),  M = 2 I + 0.3 cos(q_i - q_j),  h = 0.1 sin q,  g = 4.9 sin q
def inv_dyn_numpy(q, qd, qdd):
    q, qd, qdd = np.asarray(q), np.asarray(qd), np.asarray(qdd)
    M = 2.0 * np.eye(len(q)) + 0.3 * np.cos(q[:, None] - q[None, :])
    return M.dot(qdd) + 0.1 * np.sin(q) * np.dot(qd, qd) + 4.9 * np.sin(q)


def inv_dyn_torch(q, qd, qdd):
    """Batched form on tensors [M, dof] (same formula).  M(q) qdd is evaluated without materialising the [M, dof, dof]
    matrices: sum_j cos(q_i - q_j) qdd_j = cos q_i * sum_j cos q_j qdd_j + sin q_i * sum_j sin q_j qdd_j."""
    import torch
    cq, sq = torch.cos(q), torch.sin(q)
    mq = 2.0 * qdd + 0.3 * (cq * (cq * qdd).sum(-1, keepdim=True) + sq * (sq * qdd).sum(-1, keepdim=True))
    return mq + 0.1 * sq * (qd * qd).sum(-1, keepdim=True) + 4.9 * sq


def make_torque_problem(seed, dof=6, nway=5):
    rng = np.random.RandomState(seed)
    way = rng.randn(nway, dof)
    vl = 10 + rng.rand(dof) * 20
    al = 10 + rng.rand(dof) * 2
    tl = 40 + rng.rand(dof) * 10
    return way, np.vstack((-vl, vl)).T, np.vstack((-al, al)).T, np.vstack((-tl, tl)).T
