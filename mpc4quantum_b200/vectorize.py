"""Model construction -- API of mpc4quantum/vectorize.py.

``discretize_homogeneous`` (vectorize.py:8-49) runs in the sm_100a kernel behind ``m4q_taylor_discretize_batched``
(batched so that perturbed *models* can be discretised per ensemble member); ``vectorize_me`` (vectorize.py:52-75)
is a once-per-model host-side change of basis.
"""
import numpy as np

from . import _lib
from .linearize import create_power_list


def _as_matrix(op):
    return np.asarray(op.full() if hasattr(op, 'full') else op, dtype=complex)


def liouvillian(H):
    """Matrix of rho -> -i[H, rho] on row-major vec(rho): what vectorize_me gives in the |a><b| basis."""
    H = _as_matrix(H)
    eye = np.eye(H.shape[0])
    return -1j * (np.kron(H, eye) - np.kron(eye, H.T))


def vectorize_me(H, measure_list):
    """Liouville equation projected on an operator basis (vectorize.py:52-75).

    A[k, j] = -i * sum_{i != k} tr(H^dag s_i) * tr([s_i, s_k]^dag s_j); H and the basis may be arrays or
    objects with ``.full()`` (qutip.Qobj).
    """
    Hm = _as_matrix(H)
    basis = np.stack([_as_matrix(s) for s in measure_list])                      # [n, d, d]
    h = np.einsum('ab,iab->i', Hm.conj(), basis)                                  # tr(H^dag s_i)
    comm = np.einsum('iab,kbc->ikac', basis, basis) - np.einsum('kab,ibc->ikac', basis, basis)
    struct = np.einsum('ikab,jab->ikj', comm.conj(), basis)                       # tr([s_i, s_k]^dag s_j)
    idx = np.arange(len(basis))
    struct[idx, idx, :] = 0.0                                                     # the i == j guard (vectorize.py:60)
    return -1j * np.einsum('i,ikj->kj', h, struct)


def discretize_homogeneous_batched(L, dt, order):
    """L [N, m+1, c, c] complex (host or device) -> device tensor [N, c, c*(p+1)] (vectorize.py:8-49 per member)."""
    lib = _lib.lib()
    Ld = _lib.dev(L, np.complex128)
    n, m1, c, _ = Ld.shape
    table = np.vstack(create_power_list(order, m1 - 1)).astype(np.int32).reshape(-1, max(m1 - 1, 1))
    p1 = len(create_power_list(order, m1 - 1))
    powers = _lib.dev(table, np.int32)
    out = _lib.empty((n, c, c * p1), np.complex128)
    _lib.check(lib.m4q_taylor_discretize_batched(n, c, m1 - 1, order, p1, float(dt), _lib.ptr(Ld), _lib.ptr(powers),
                                                 _lib.ptr(out), _lib.stream_ptr()))
    return out


def discretize_homogeneous(A_cts_list, dt, order):
    """Order-`order` Taylor discretisation of exp((A_0 + sum u_i A_i) dt) grouped by control monomial.

    Same arguments and return layout as vectorize.py:8-49: hstack of C(order+m, m) blocks in create_power_list order.
    """
    L = np.stack([np.asarray(a, dtype=complex) for a in A_cts_list])[None]
    return discretize_homogeneous_batched(L, dt, order)[0].cpu().numpy()
