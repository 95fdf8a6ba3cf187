"""fp64 CPU restatement of the reference ``QuantumConv1d`` (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/quantum_whisper.py``:
  * ``:47-59,88``  parameters (``pre_conv``, ``post_conv``, ``quantum_weights``)
  * ``:64-85``     the circuit (pad, AmplitudeEmbedding(normalize=True), Rot per wire,
                   CNOT chain, <PauliZ(i)>)
  * ``:95-128``    padding, window extraction, loop order, fp32 cast of the readout

The simulator arithmetic lives in the un-vendored third-party dependency
``pennylane>=0.30.0`` (``requirements.txt:12``, unpinned).  Its documented conventions are
restated here:
  * wire 0 is the MOST significant bit of the basis-state index;
  * ``RZ(a) = diag(e^{-ia/2}, e^{+ia/2})``, ``RY(a) = [[c,-s],[s,c]]``;
  * ``Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi)``;
  * ``AmplitudeEmbedding(normalize=True)`` divides by ``sqrt(sum |v|^2)`` inside autograd;
  * ``CNOT(control, target)``; ``expval(PauliZ(i)) = sum_k |psi_k|^2 (1 - 2 bit_i(k))``;
  * gradients = exact analytic derivatives (backprop through the simulator).

PARITY UNPINNED: no PennyLane here, no reference test for this path (see
``oracle/__init__.py``).  Everything is done in float64 / complex128 and cast to float32
exactly where ``quantum_whisper.py:122`` does (``.float()``), when ``cast_fp32=True``.

Extensions that do not exist in the reference (their semantics are defined HERE and the
CUDA kernels are checked against this file):
  * ``n_layers = Lq > 1``: the block ``[Rot on every wire; CNOT chain]`` is repeated ``Lq``
    times with ``quantum_weights`` of shape ``(Lq, q, 3)``.
  * ``embedding = "angle"``: replaces pad+AmplitudeEmbedding by, on each wire ``i`` of the
    all-zeros state, ``RY(pre_i)`` followed by ``RZ(pre_i)``.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

EMB_AMPLITUDE = 0
EMB_ANGLE = 1


# ----------------------------------------------------------------------------- gates
def _rot_matrix(phi: torch.Tensor, theta: torch.Tensor, omega: torch.Tensor) -> torch.Tensor:
    """2x2 complex128 matrix of Rot(phi,theta,omega)=RZ(omega)RY(theta)RZ(phi).

    (quantum_whisper.py:78; matrix as in SURVEY.md 8-a6.)
    """
    c = torch.cos(theta / 2)
    s = torch.sin(theta / 2)
    ep = torch.exp(-0.5j * (phi + omega))
    em = torch.exp(0.5j * (phi - omega))
    row0 = torch.stack([ep * c, -em * s])
    row1 = torch.stack([torch.conj(em) * s, torch.conj(ep) * c])
    return torch.stack([row0, row1])


def _apply_1q(state: torch.Tensor, gate: torch.Tensor, wire: int) -> torch.Tensor:
    """state: (W, 2, ..., 2) with wire ``w`` on axis ``1+w`` (big-endian). gate: (2,2) or (W,2,2)."""
    st = torch.movedim(state, 1 + wire, -1)  # (..., 2)
    if gate.dim() == 2:
        st = torch.einsum("...k,jk->...j", st, gate)
    else:
        shp = st.shape
        st = torch.einsum("wnk,wjk->wnj", st.reshape(shp[0], -1, 2), gate).reshape(shp)
    return torch.movedim(st, -1, 1 + wire)


def _apply_cnot(state: torch.Tensor, control: int, target: int) -> torch.Tensor:
    """CNOT(control,target): flip the target axis on the control=1 slice (quantum_whisper.py:82)."""
    idx0 = [slice(None)] * state.dim()
    idx1 = [slice(None)] * state.dim()
    idx0[1 + control] = slice(0, 1)
    idx1[1 + control] = slice(1, 2)
    s0 = state[tuple(idx0)]
    s1 = torch.flip(state[tuple(idx1)], dims=[1 + target])
    return torch.cat([s0, s1], dim=1 + control)


def embed(pre: torch.Tensor, q: int, embedding: int) -> torch.Tensor:
    """(W,q) float64 -> (W,2,...,2) complex128 initial state."""
    W = pre.shape[0]
    N = 1 << q
    if embedding == EMB_AMPLITUDE:
        # quantum_whisper.py:67-74: zero-pad to 2^q (always the pad branch since q < 2^q),
        # then normalise; ||v|| = 0 gives NaN exactly as the reference (not guarded).
        v = torch.cat([pre, torch.zeros(W, N - q, dtype=pre.dtype)], dim=1)
        v = v / torch.sqrt(torch.sum(v * v, dim=1, keepdim=True))
        return v.to(torch.complex128).reshape((W,) + (2,) * q)
    elif embedding == EMB_ANGLE:
        st = torch.zeros(W, N, dtype=torch.complex128)
        st[:, 0] = 1.0
        st = st.reshape((W,) + (2,) * q)
        for i in range(q):
            a = pre[:, i]
            c = torch.cos(a / 2).to(torch.complex128)
            s = torch.sin(a / 2).to(torch.complex128)
            ry = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)  # (W,2,2)
            st = _apply_1q(st, ry, i)
            e = torch.exp(-0.5j * a.to(torch.complex128))
            zero = torch.zeros_like(e)
            rz = torch.stack([torch.stack([e, zero], -1), torch.stack([zero, torch.conj(e)], -1)], -2)
            st = _apply_1q(st, rz, i)
        return st
    raise ValueError(f"unknown embedding {embedding}")


def circuit_expvals(pre: torch.Tensor, qweights: torch.Tensor, embedding: int = EMB_AMPLITUDE) -> torch.Tensor:
    """Vectorised statevector restatement of the QNode (quantum_whisper.py:64-85).

    pre: (W,q) float64; qweights: (q,3) or (Lq,q,3) float64.  Returns (W,q) float64 <Z_i>.
    Differentiable (torch autograd == the reference's backprop through default.qubit).
    """
    assert pre.dtype == torch.float64 and qweights.dtype == torch.float64
    W, q = pre.shape
    qw = qweights if qweights.dim() == 3 else qweights[None]
    assert qw.shape[1:] == (q, 3)
    st = embed(pre, q, embedding)
    for layer in range(qw.shape[0]):
        for i in range(q):  # :77-78
            st = _apply_1q(st, _rot_matrix(qw[layer, i, 0], qw[layer, i, 1], qw[layer, i, 2]), i)
        for i in range(q - 1):  # :81-82
            st = _apply_cnot(st, i, i + 1)
    prob = (st.real**2 + st.imag**2)  # (W,2,...,2)
    outs = []
    for i in range(q):  # :85
        dims = [d for d in range(1, q + 1) if d != 1 + i]
        marg = prob.sum(dim=dims) if dims else prob
        outs.append(marg[:, 0] - marg[:, 1])
    return torch.stack(outs, dim=1)


def circuit_single_window(pre_vec: torch.Tensor, qweights: torch.Tensor) -> torch.Tensor:
    """One window, written out gate by gate on a flat 2^q complex vector -- the literal shape of
    one reference QNode call (quantum_whisper.py:64-85), used by the literal-loop CPU baseline."""
    q = pre_vec.shape[0]
    N = 1 << q
    padded = torch.cat([pre_vec, torch.zeros(N - q, dtype=pre_vec.dtype)])  # :69
    psi = (padded / torch.linalg.vector_norm(padded)).to(torch.complex128)  # :74
    psi = psi.reshape((2,) * q)
    for i in range(q):  # :77-78
        g = _rot_matrix(qweights[i, 0], qweights[i, 1], qweights[i, 2])
        psi = torch.movedim(torch.tensordot(g, psi, dims=([1], [i])), 0, i)
    for i in range(q - 1):  # :81-82
        idx = [slice(None)] * q
        idx[i] = 1
        sub = psi[tuple(idx)]
        flipped = torch.flip(sub, dims=[i])  # target axis i+1 becomes axis i after indexing out axis i
        psi = torch.stack([psi[tuple(idx[:i] + [0] + idx[i + 1:])], flipped], dim=i)
    prob = psi.real**2 + psi.imag**2
    outs = []
    for i in range(q):  # :85
        dims = [d for d in range(q) if d != i]
        m = prob.sum(dim=dims) if dims else prob
        outs.append(m[0] - m[1])
    return torch.stack(outs)


# ------------------------------------------------- second, independent oracle (dense unitary)
def dense_unitary(qweights: np.ndarray) -> np.ndarray:
    """Full 2^q x 2^q unitary of the trainable part, built with Kronecker products and an explicit
    CNOT permutation matrix (independent of the tensor-axis code above). numpy complex128."""
    qw = np.asarray(qweights, dtype=np.float64)
    if qw.ndim == 2:
        qw = qw[None]
    Lq, q, _ = qw.shape
    N = 1 << q
    U = np.eye(N, dtype=np.complex128)
    for layer in range(Lq):
        G = np.ones((1, 1), dtype=np.complex128)
        for i in range(q):
            phi, th, om = qw[layer, i]
            rz1 = np.diag([np.exp(-0.5j * phi), np.exp(0.5j * phi)])
            ry = np.array([[math.cos(th / 2), -math.sin(th / 2)], [math.sin(th / 2), math.cos(th / 2)]])
            rz2 = np.diag([np.exp(-0.5j * om), np.exp(0.5j * om)])
            G = np.kron(G, rz2 @ ry @ rz1)  # wire 0 leftmost = most significant
        U = G @ U
        for i in range(q - 1):
            Pm = np.zeros((N, N))
            for k in range(N):
                cbit = (k >> (q - 1 - i)) & 1
                k2 = k ^ (cbit << (q - 1 - (i + 1)))
                Pm[k2, k] = 1.0
            U = Pm @ U
    return U


def dense_unitary_expvals(pre: np.ndarray, qweights: np.ndarray) -> np.ndarray:
    """out_i = xh^T M_i xh with M_i = Re(U[:, :q]^H Z_i U[:, :q]) (structure (iii))."""
    pre = np.asarray(pre, dtype=np.float64)
    W, q = pre.shape
    N = 1 << q
    U = dense_unitary(qweights)[:, :q]
    xh = pre / np.linalg.norm(pre, axis=1, keepdims=True)
    out = np.empty((W, q))
    ks = np.arange(N)
    for i in range(q):
        z = 1.0 - 2.0 * ((ks >> (q - 1 - i)) & 1)
        M = np.real(U.conj().T @ (z[:, None] * U))
        out[:, i] = np.einsum("wa,ab,wb->w", xh, M, xh)
    return out


# ----------------------------------------------------------------------------- windowing
def out_length(L: int, K: int, S: int, P: int) -> int:
    """quantum_whisper.py:103."""
    return (L + 2 * P - K) // S + 1


def window_indices(L: int, K: int, S: int, P: int) -> np.ndarray:
    """(L_out, K) int64 ORIGINAL column index touched by (window i, tap k); -1 == zero padding.

    quantum_whisper.py:99-110: pad P zeros each side, window i = padded columns [i*S, i*S+K).
    """
    Lo = out_length(L, K, S, P)
    idx = np.arange(Lo)[:, None] * S - P + np.arange(K)[None, :]
    idx[(idx < 0) | (idx >= L)] = -1
    return idx


def extract_windows(x: torch.Tensor, K: int, S: int, P: int) -> torch.Tensor:
    """(B,C,L) -> (B,L_out,C*K), feature index f = c*K + k (quantum_whisper.py:108-111)."""
    B, C, L = x.shape
    xp = torch.nn.functional.pad(x, (P, P)) if P > 0 else x
    Lo = out_length(L, K, S, P)
    win = xp.unfold(2, K, S)[:, :, :Lo]  # (B,C,Lo,K)
    return win.permute(0, 2, 1, 3).reshape(B, Lo, C * K)


# ----------------------------------------------------------------------------- the layer
def qconv1d_forward(
    x: torch.Tensor,
    w_pre: torch.Tensor,
    b_pre: torch.Tensor,
    qweights: torch.Tensor,
    w_post: torch.Tensor,
    b_post: torch.Tensor,
    K: int,
    S: int = 1,
    P: int = 0,
    embedding: int = EMB_AMPLITUDE,
    cast_fp32: bool = False,
    return_intermediates: bool = False,
):
    """Vectorised restatement of QuantumConv1d.forward (quantum_whisper.py:95-128). All float64.

    ``cast_fp32=True`` rounds the readout to float32 where the reference does (:122).
    """
    B, C, L = x.shape
    q = w_pre.shape[0]
    win = extract_windows(x, K, S, P)  # (B,Lo,CK)
    Lo = win.shape[1]
    pre = win @ w_pre.T + b_pre  # :114
    qout = circuit_expvals(pre.reshape(B * Lo, q), qweights, embedding).reshape(B, Lo, q)
    if cast_fp32:
        qout = qout.float().double()  # :122
    y = qout @ w_post.T + b_post  # :125
    y = y.permute(0, 2, 1).contiguous()  # :126 column i of (B,O,L_out)
    if return_intermediates:
        return y, pre, qout
    return y


def stem_forward(x, params1, params2, positional_embedding=None):
    """fp64 restatement of the stem of the quantum audio encoder: whisper/whisper/model.py:193-198 with conv1/conv2 the
    QuantumConv1d layers of quantum_whisper.py:136-137 (k3/s1/p1 then k3/s2/p1); exact-erf GELU (torch default).
    x (B, n_mels, L) -> (B, L_out2, n_state)."""
    h = torch.nn.functional.gelu(qconv1d_forward(x, *params1, K=3, S=1, P=1))
    h = torch.nn.functional.gelu(qconv1d_forward(h, *params2, K=3, S=2, P=1))
    h = h.permute(0, 2, 1)
    if positional_embedding is not None:
        assert tuple(h.shape[1:]) == tuple(positional_embedding.shape), "incorrect audio shape"  # model.py:197
        h = h + positional_embedding
    return h


def qconv1d_literal(x, w_pre, b_pre, qweights, w_post, b_post, K, S=1, P=0, max_windows: Optional[int] = None):
    """Literal loop nest of the reference forward (quantum_whisper.py:107-126): for each output
    column, for each batch element, one single-window simulation.  float64.  Used as the
    "restatement of reference -- PennyLane unavailable" CPU baseline.
    ``max_windows`` bounds the number of output columns processed (the rest stay zero)."""
    B, C, L = x.shape
    O = w_post.shape[0]
    if P > 0:
        x = torch.nn.functional.pad(x, (P, P))
    Lo = out_length(L, K, S, P)
    out = torch.zeros(B, O, Lo, dtype=x.dtype)
    n = Lo if max_windows is None else min(Lo, max_windows)
    cols = []
    for i in range(n):
        s = i * S
        flat = x[:, :, s:s + K].reshape(B, -1)
        pre = torch.nn.functional.linear(flat, w_pre, b_pre)
        qs = []
        for j in range(B):
            qs.append(circuit_single_window(pre[j], qweights))
        qo = torch.stack(qs)
        cols.append(torch.nn.functional.linear(qo, w_post, b_post))
    if cols:
        out = torch.cat([torch.stack(cols, dim=2), out[:, :, n:]], dim=2)
    return out


def qconv1d_grads(x, params, gy, K, S, P, embedding=EMB_AMPLITUDE, need_gx=True):
    """Backward of the layer by autograd through the fp64 restatement (== the reference's
    backprop).  Returns dict of float64 grads."""
    leaves = [t.detach().clone().requires_grad_(True) for t in params]
    xin = x.detach().clone().requires_grad_(need_gx)
    y = qconv1d_forward(xin, *leaves, K=K, S=S, P=P, embedding=embedding)
    wanted = leaves + ([xin] if need_gx else [])
    g = torch.autograd.grad(y, wanted, grad_outputs=gy)
    names = ["w_pre", "b_pre", "qweights", "w_post", "b_post"] + (["x"] if need_gx else [])
    res = dict(zip(names, g))
    res["y"] = y.detach()
    return res


def make_params(C: int, O: int, K: int, q: int, n_layers: int = 1, seed: int = 0, dtype=torch.float64):
    """Deterministic parameters with the reference's shapes (quantum_whisper.py:58-59,88) and the
    default nn.Linear init distribution (U(-1/sqrt(fan_in), 1/sqrt(fan_in)))."""
    g = torch.Generator().manual_seed(seed)
    q = min(q, C * K)
    kpre = 1.0 / math.sqrt(C * K)
    kpost = 1.0 / math.sqrt(q)
    w_pre = (torch.rand(q, C * K, generator=g, dtype=torch.float64) * 2 - 1) * kpre
    b_pre = (torch.rand(q, generator=g, dtype=torch.float64) * 2 - 1) * kpre
    shape = (q, 3) if n_layers == 1 else (n_layers, q, 3)
    qw = torch.randn(*shape, generator=g, dtype=torch.float64)
    w_post = (torch.rand(O, q, generator=g, dtype=torch.float64) * 2 - 1) * kpost
    b_post = (torch.rand(O, generator=g, dtype=torch.float64) * 2 - 1) * kpost
    return tuple(t.to(dtype) for t in (w_pre, b_pre, qw, w_post, b_post))
