"""CPU oracle for the OQRL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``oqrl_b200``) never imports it and has no CPU fallback.

What it restates (float64 numpy, inputs rounded to float32 first):

* the circuit of ``/root/reference/src/agent.py:36-46`` -- AngleEmbedding(RY) +
  StronglyEntanglingLayers + <Z_i> readout.  The arithmetic of those two
  templates lives in PennyLane (``requirements.txt:7`` ``pennylane>=0.32.0``,
  unpinned, NOT vendored, NOT installable here), so the published template
  definitions are restated:
    - wire 0 is the most significant bit of the basis index,
    - RY(t) = [[c,-s],[s,c]], c=cos(t/2), s=sin(t/2); RZ(t)=diag(e^{-it/2},e^{+it/2}),
    - Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi),
    - layer l: Rot on every wire, then (n>1) CNOT(i -> (i+r_l) mod n) for
      i=0..n-1 in order, r_l = (l mod (n-1)) + 1,
    - parameter index (l*n + i)*3 + {0:phi,1:theta,2:omega} (agent.py:39-43).
* the batch semantics of ``agent.py:48-61`` (one circuit per row),
* the fitness of ``/root/reference/src/optim.py:152-173``:
      fitness = -mean_b (Q(s_b)[a_b] - (r_b + (1-d_b)*0.99*max_a Q(s'_b)[a]))^2
  with BOTH Q and Q' evaluated with the candidate's own weights.

PARITY STATUS: **parity unpinned by the reference** -- the reference has no
test, fixture or golden vector for this path and PennyLane cannot be run here.
The pins are (a) two structurally independent evaluators in this file
(`qvalues_tensor`: per-gate axis contraction; `qvalues_dense`: dense 2^n x 2^n
Kronecker unitaries with projector-built CNOTs) that must agree to 1e-12,
(b) hand-derivable basis-state cases that fix CNOT direction and the range
schedule, and (c) the survey-derived known-answer vectors of SURVEY.md
Appendix B (tests/golden/kat_appendix_b.json).

Builder-defined generalisations (no reference counterpart):
``reupload`` (re-apply the embedding before every layer), ``feature_map=tiled``
(RY(x[i mod n_feat]) on every wire) and ``entangler=cz_ring``
(CZ(i,(i+1) mod n) for i=0..n-1; for n==2 that is CZ(0,1) twice = identity).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEL_CNOT = 0
CZ_RING = 1
FEAT_FIRST_K = 0
FEAT_TILED = 1


@dataclass(frozen=True)
class Spec:
    """Circuit description; mirrors ``oqrl_spec`` in include/oqrl_b200.h."""

    n_qubits: int = 4
    n_layers: int = 2
    n_out: int = 2
    n_feat: int = 4
    entangler: int = SEL_CNOT
    reupload: int = 0
    feature_map: int = FEAT_FIRST_K

    @property
    def n_params(self) -> int:
        return self.n_layers * self.n_qubits * 3

    def validate(self) -> None:
        n = self.n_qubits
        if not (1 <= n <= 30) or self.n_layers < 0:
            raise ValueError("bad n_qubits/n_layers")
        if not (1 <= self.n_out <= n):
            raise ValueError("n_out must be in 1..n_qubits")
        if self.n_feat < 0:
            raise ValueError("n_feat < 0")
        if self.feature_map == FEAT_FIRST_K and self.n_feat > n:
            # AngleEmbedding raises when there are more features than wires.
            raise ValueError("more features than wires")


def sel_range(n: int, layer: int) -> int:
    """Default StronglyEntanglingLayers range r_l = (l mod (n-1)) + 1."""
    return (layer % (n - 1)) + 1 if n > 1 else 0


def embed_wires(spec: Spec):
    """[(wire, feature index)] rotated by the embedding."""
    if spec.n_feat == 0:
        return []
    if spec.feature_map == FEAT_TILED:
        return [(w, w % spec.n_feat) for w in range(spec.n_qubits)]
    return [(w, w) for w in range(spec.n_feat)]


# ----------------------------------------------------------------------------
# gate matrices
# ----------------------------------------------------------------------------
def ry_matrix(t):
    """RY(t) for an array of angles -> [..., 2, 2] float64."""
    t = np.asarray(t, dtype=np.float64)
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)


def rot_matrix(phi, theta, omega):
    """Rot(phi,theta,omega)=RZ(omega)RY(theta)RZ(phi) -> [..., 2, 2] complex128."""
    phi, theta, omega = (np.asarray(a, dtype=np.float64) for a in (phi, theta, omega))
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    m00 = np.exp(-0.5j * (phi + omega)) * c
    m01 = -np.exp(0.5j * (phi - omega)) * s
    m10 = np.exp(-0.5j * (phi - omega)) * s
    m11 = np.exp(0.5j * (phi + omega)) * c
    return np.stack([np.stack([m00, m01], -1), np.stack([m10, m11], -1)], -2)


# ----------------------------------------------------------------------------
# evaluator 1: per-gate axis contraction, vectorised over a leading batch
# ----------------------------------------------------------------------------
def _apply_1q(state, mats, wire, n):
    """state [N, 2^n] complex; mats [N,2,2] or [2,2]; gate on `wire`."""
    N = state.shape[0]
    v = state.reshape(N, 1 << wire, 2, 1 << (n - 1 - wire))
    if mats.ndim == 2:
        out = np.einsum("ab,nxby->nxay", mats, v)
    else:
        out = np.einsum("nab,nxby->nxay", mats, v)
    return out.reshape(N, 1 << n)


def _bit(k, wire, n):
    return (k >> (n - 1 - wire)) & 1


def _apply_cnot(state, control, target, n):
    k = np.arange(1 << n)
    src = k ^ (_bit(k, control, n) << (n - 1 - target))
    return state[:, src]


def _apply_cz(state, a, b, n):
    k = np.arange(1 << n)
    sign = 1.0 - 2.0 * (_bit(k, a, n) & _bit(k, b, n))
    return state * sign


def simulate(spec: Spec, params, x):
    """Final statevectors.  params [D] or [N,D]; x [N,n_feat] -> [N, 2^n] complex128."""
    spec.validate()
    n, L = spec.n_qubits, spec.n_layers
    x = np.asarray(np.asarray(x, dtype=np.float32), dtype=np.float64)
    if x.ndim == 1:
        x = x[None]
    N = x.shape[0]
    params = np.asarray(np.asarray(params, dtype=np.float32), dtype=np.float64)
    per_sample = params.ndim == 2
    w = params.reshape((N, L, n, 3)) if per_sample else params.reshape((L, n, 3))
    state = np.zeros((N, 1 << n), dtype=np.complex128)
    state[:, 0] = 1.0
    emb = embed_wires(spec)

    def embed(st):
        for wire, f in emb:
            st = _apply_1q(st, ry_matrix(x[:, f]).astype(np.complex128), wire, n)
        return st

    if not spec.reupload or L == 0:
        state = embed(state)
    for l in range(L):
        if spec.reupload:
            state = embed(state)
        for i in range(n):
            a = w[:, l, i] if per_sample else w[l, i]
            state = _apply_1q(state, rot_matrix(a[..., 0], a[..., 1], a[..., 2]), i, n)
        if n > 1:
            if spec.entangler == SEL_CNOT:
                r = sel_range(n, l)
                for i in range(n):
                    state = _apply_cnot(state, i, (i + r) % n, n)
            else:
                for i in range(n):
                    state = _apply_cz(state, i, (i + 1) % n, n)
    return state


def expval_z(state, n, n_out):
    """<Z_i> for i<n_out from [N,2^n] amplitudes -> [N,n_out] float64."""
    k = np.arange(1 << n)
    prob = (state.real**2 + state.imag**2)
    out = np.empty((state.shape[0], n_out))
    for i in range(n_out):
        out[:, i] = prob @ (1.0 - 2.0 * _bit(k, i, n))
    return out


def qvalues_tensor(spec: Spec, params, x):
    """Q-values [N, n_out] float64 (agent.py:48-61 semantics, one circuit per row)."""
    return expval_z(simulate(spec, params, x), spec.n_qubits, spec.n_out)


# ----------------------------------------------------------------------------
# evaluator 2: dense unitaries (independent construction, small n only)
# ----------------------------------------------------------------------------
def _kron_on(mat, wire, n):
    out = np.array([[1.0 + 0j]])
    for q in range(n):
        out = np.kron(out, mat if q == wire else np.eye(2))
    return out


def _dense_cnot(c, t, n):
    p0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
    p1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
    xg = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    a = np.array([[1.0 + 0j]])
    b = np.array([[1.0 + 0j]])
    for q in range(n):
        a = np.kron(a, p0 if q == c else np.eye(2))
        b = np.kron(b, p1 if q == c else (xg if q == t else np.eye(2)))
    return a + b


def _dense_cz(c, t, n):
    p1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
    b = np.array([[1.0 + 0j]])
    for q in range(n):
        b = np.kron(b, p1 if q in (c, t) else np.eye(2))
    return np.eye(1 << n) - 2 * b


def qvalues_dense(spec: Spec, params, x):
    """Same contract as qvalues_tensor, built from full 2^n x 2^n matrices."""
    spec.validate()
    n, L = spec.n_qubits, spec.n_layers
    assert n <= 8, "dense evaluator is for small n"
    x = np.asarray(np.asarray(x, dtype=np.float32), dtype=np.float64)
    if x.ndim == 1:
        x = x[None]
    params = np.asarray(np.asarray(params, dtype=np.float32), dtype=np.float64)
    zdiag = [np.real(np.diag(_kron_on(np.diag([1.0, -1.0]).astype(complex), i, n))) for i in range(spec.n_out)]
    out = np.empty((x.shape[0], spec.n_out))
    emb = embed_wires(spec)
    for b in range(x.shape[0]):
        w = (params[b] if params.ndim == 2 else params).reshape(L, n, 3)
        U = np.eye(1 << n, dtype=np.complex128)

        def embed(U):
            for wire, f in emb:
                U = _kron_on(ry_matrix(x[b, f]).astype(complex), wire, n) @ U
            return U

        if not spec.reupload or L == 0:
            U = embed(U)
        for l in range(L):
            if spec.reupload:
                U = embed(U)
            for i in range(n):
                U = _kron_on(rot_matrix(*w[l, i]), i, n) @ U
            if n > 1:
                if spec.entangler == SEL_CNOT:
                    r = sel_range(n, l)
                    for i in range(n):
                        U = _dense_cnot(i, (i + r) % n, n) @ U
                else:
                    for i in range(n):
                        U = _dense_cz(i, (i + 1) % n, n) @ U
        psi = U[:, 0]
        p = np.abs(psi) ** 2
        out[b] = [p @ z for z in zdiag]
    return out


# ----------------------------------------------------------------------------
# fitness (optim.py:152-173) and ranking (optim.py:207)
# ----------------------------------------------------------------------------
GAMMA_REF = 0.99  # literal at optim.py:163


def td_fitness_from_q(q, q_next, act, rew, term, gamma=GAMMA_REF):
    """q,q_next [P,T,n_out]; act [T] int; rew,term [T] -> fitness [P] float64."""
    act = np.asarray(act).reshape(-1).astype(np.int64)
    rew = np.asarray(np.asarray(rew, dtype=np.float32), dtype=np.float64).reshape(-1)
    term = np.asarray(np.asarray(term, dtype=np.float32), dtype=np.float64).reshape(-1)
    qa = np.take_along_axis(q, act[None, :, None], axis=2)[..., 0]
    # optim.py:163: `(1 - terminals) * 0.99` is a float32 tensor times a python scalar, i.e. it is
    # evaluated in float32 (0.99 -> 0.99000001) BEFORE it meets the float64 Q-values.
    keep = ((np.float32(1.0) - term.astype(np.float32)) * np.float32(gamma)).astype(np.float64)
    tgt = rew[None] + keep[None] * q_next.max(axis=2)
    return -np.mean((qa - tgt) ** 2, axis=1)


def fitness(spec: Spec, params, obs, act, rew, next_obs, term, gamma=GAMMA_REF, chunk=1 << 16):
    """Population fitness.  params [P,D]; obs,next_obs [T,n_feat] -> [P] float64."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float32))
    obs = np.asarray(obs, dtype=np.float32)
    next_obs = np.asarray(next_obs, dtype=np.float32)
    P, T = params.shape[0], obs.shape[0]
    q = np.empty((P, T, spec.n_out))
    qn = np.empty((P, T, spec.n_out))
    rows = max(1, chunk // max(T, 1))
    for p0 in range(0, P, rows):
        p1 = min(P, p0 + rows)
        pp = np.repeat(params[p0:p1], T, axis=0)
        q[p0:p1] = qvalues_tensor(spec, pp, np.tile(obs, (p1 - p0, 1))).reshape(p1 - p0, T, -1)
        qn[p0:p1] = qvalues_tensor(spec, pp, np.tile(next_obs, (p1 - p0, 1))).reshape(p1 - p0, T, -1)
    return td_fitness_from_q(q, qn, act, rew, term, gamma)


def ranking(fit):
    """np.argsort(fitness)[::-1] exactly as optim.py:207 does it."""
    return np.argsort(np.asarray(fit))[::-1]


def flops_per_eval(spec: Spec) -> int:
    """Canonical gate-by-gate flop count F."""
    n, L = spec.n_qubits, spec.n_layers
    n_emb = len(embed_wires(spec))
    reps = L if (spec.reupload and L > 0) else 1
    return (1 << n) * (6 * n_emb * reps + 14 * n * L + (3 + spec.n_out))
