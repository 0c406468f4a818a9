"""Gauss-Legendre rules on the mesh spans (host side; the tables are uploaded once).

Restates ``pyiga/quadrature.py:3-21``: ``leggauss(nqp)`` mapped affinely to every span,
``node = h*x + m``, ``weight = h*w`` with ``m``/``h`` the span mid point / half length.
"""
import numpy as np


def gauss_rule(deg, a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    mid, half = 0.5 * (a + b), 0.5 * (b - a)
    x, w = np.polynomial.legendre.leggauss(deg)
    return (np.outer(half, x) + mid[:, None]).ravel(), np.outer(half, w).ravel()


def make_iterated_quadrature(intervals, nqp):
    intervals = np.asarray(intervals, dtype=float)
    return gauss_rule(nqp, intervals[:-1], intervals[1:])


def make_tensor_quadrature(meshes, nqp):
    rules = [make_iterated_quadrature(mesh, nqp) for mesh in meshes]
    return tuple(r[0] for r in rules), tuple(r[1] for r in rules)
