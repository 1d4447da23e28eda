"""Numerical evaluation of a CompiledForm's pointwise expressions (test helper)."""
import numpy as np
import sympy as sp

from echemfem_b200 import symbolic as S


def random_point_data(nf, naux, d, npts, rng):
    def unit(v):
        return v / np.linalg.norm(v, axis=-1, keepdims=True)
    dat = {"x": rng.random((npts, d)), "n": unit(rng.normal(size=(npts, d))),
           "u": 0.5 + rng.random((nf, npts)), "gu": rng.normal(size=(nf, npts, d)),
           "um": 0.5 + rng.random((nf, npts)), "gum": rng.normal(size=(nf, npts, d)),
           "a": rng.normal(size=(max(naux, 1), npts)), "am": rng.normal(size=(max(naux, 1), npts)),
           "cv": 0.1 + rng.random(npts), "cvm": 0.1 + rng.random(npts),
           "cd": 0.5 + rng.random(npts), "fa": 0.1 + rng.random(npts)}
    return dat


def _value(name, dat):
    base, side = (name[:-2], "m") if name.endswith("_m") else ((name[:-2], "") if name.endswith("_p") else (name, ""))
    if base[0] == "x" and base[1:].isdigit():
        return dat["x"][:, int(base[1:])]
    if base[0] == "n" and base[1:].isdigit():
        return dat["n"][:, int(base[1:])]
    if base in ("cv", "cd"):
        return dat.get(base + side, dat[base])
    if base == "fa":
        return dat["fa"]
    if base[0] == "k" and base[1:].isdigit():
        return np.full(dat["x"].shape[0], S.Constant.lookup(sp.Symbol(base)))
    kind, j, comp, _ = S.parse_field_symbol(sp.Symbol(base))
    if comp is None:
        return dat[kind + side][j]
    return dat["g" + kind + side][j][:, comp]


def evaluate(expr, dat, form=None):
    expr = sp.sympify(expr)
    if form is not None:
        expr = expr.xreplace({sp.Symbol(b, real=True): sp.Symbol(a, real=True)
                              for a, b in zip(form.const_source, form.const_names)})
    syms = sorted(expr.free_symbols, key=str)
    npts = dat["x"].shape[0]
    if not syms:
        return np.full(npts, float(expr))
    f = sp.lambdify(syms, expr, [{"rpow": np.power, "PosPart": lambda v: np.maximum(v, 0.0)}, "numpy"])
    return np.broadcast_to(f(*[_value(str(s), dat) for s in syms]), (npts,)).astype(float)
