"""Import the UNMODIFIED reference package from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference needs qutip,
cvxpy and matplotlib, none of which exist in this image, and two numpy names
removed in numpy 2 (``np.product`` linearize.py:126 / vectorize.py:44,
``np.math`` vectorize.py:32).  This module installs the smallest possible
stand-ins so that the reference's own pure-numpy code runs verbatim:

* ``mpc4quantum.mpc.mpc`` (mpc.py:128-304), ``StepClock``, ``iqp_line_search``
* ``mpc4quantum.linearize.*`` (WrapModel, library helpers, krtimes)
* ``mpc4quantum.vectorize.discretize_homogeneous`` (vectorize.py:8-49)
* ``mpc4quantum.vectorize.vectorize_me`` (vectorize.py:52-75) through a tiny
  ``Qobj`` stand-in that implements only ``*``, ``.dag()``, ``.tr()`` and
  ``qutip.commutator``
* ``mpc4quantum.model.DMDc`` and ``QCoupledExperiment.lift/proj``

The two leaves that live in absent third-party packages are injected by the
caller: ``inject_qp(fn)`` replaces ``quad_program`` as seen by ``mpc.py:189``,
and the plant is any ``Experiment`` subclass with a numpy ``simulate``.

The reference tree does not exist on the GPU box: nothing that runs there may
import this module (``available()`` returns False there).
"""
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'mpc4quantum'))


class _MiniQobj:
    """Just enough of qutip.Qobj for vectorize.py:52-75 (``*``, ``dag``, ``tr``, ``full``)."""

    def __init__(self, data):
        self.data = np.array(data, dtype=complex)

    def __mul__(self, other):
        if isinstance(other, _MiniQobj):
            return _MiniQobj(self.data @ other.data)
        return _MiniQobj(self.data * other)

    __rmul__ = lambda self, k: _MiniQobj(self.data * k)

    def __sub__(self, other):
        return _MiniQobj(self.data - other.data)

    def __add__(self, other):
        return _MiniQobj(self.data + other.data)

    def dag(self):
        return _MiniQobj(self.data.conj().T)

    def tr(self):
        return complex(np.trace(self.data))

    def full(self):
        return self.data

    @property
    def shape(self):
        return self.data.shape


def _install_stubs():
    if not hasattr(np, 'product'):
        np.product = np.prod
    if not hasattr(np, 'math'):
        np.math = math
    if 'qutip' not in sys.modules:
        qt = types.ModuleType('qutip')
        qt.Qobj = _MiniQobj
        qt.commutator = lambda a, b: a * b - b * a
        qt.mesolve = None
        qt.propagator = None
        qt.tensor = None
        sys.modules['qutip'] = qt
    for name in ('cvxpy', 'matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


_loaded = None


def load():
    """Return the reference package object (``import mpc4quantum`` from /root/reference)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree %s is not present (GPU box?)' % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mpc4quantum  # noqa: the reference itself
    _loaded = mpc4quantum
    return _loaded


def module(name):
    """``module('mpc')`` -> the reference's mpc.py module object (m4q.mpc is the *function*)."""
    load()
    return sys.modules['mpc4quantum.' + name]


def inject_qp(fn):
    """Replace the cvxpy/OSQP leaf as seen by the reference's mpc loop (mpc.py:2, :189)."""
    module('mpc').quad_program = fn


def Qobj(a):
    load()
    return _MiniQobj(a)
