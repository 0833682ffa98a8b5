"""Generates tests/golden/*.npz by running the parts of the reference that are
importable in this container (pure NumPy: the two optimizers and the dict<->array
converters) and by reading the reference's only shipped fixture (the linear
regression CSVs).  Run once, in the build container, where /root/reference is
mounted; the outputs are committed, /root/reference is never read at test time.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load_ref_optimizers():
    # stein/optimizers/*.py use relative imports -> load them under an alias
    # package so they never shadow this repo's own `stein` package.
    pkg = types.ModuleType("_ref_stein_optimizers")
    pkg.__path__ = [os.path.join(REF, "stein", "optimizers")]
    sys.modules[pkg.__name__] = pkg
    mods = {}
    for name in ("abstract_gradient_descent", "adam_gradient_descent", "adagrad_gradient_descent"):
        spec = importlib.util.spec_from_file_location(
            pkg.__name__ + "." + name, os.path.join(pkg.__path__[0], name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["adam_gradient_descent"].AdamGradientDescent, \
        mods["adagrad_gradient_descent"].AdagradGradientDescent


def _load_ref_converters():
    spec = importlib.util.spec_from_file_location(
        "_ref_converters", os.path.join(REF, "stein", "utilities", "converters.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class FakeVar:
    """Duck-typed stand-in for a tf.Variable: `.name`, `.get_shape().as_list()`."""

    def __init__(self, name, shape):
        self.name, self._shape = name, list(shape)

    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self._shape))


def main():
    Adam, Adagrad = _load_ref_optimizers()
    conv = _load_ref_converters()
    rng = np.random.default_rng(20261018)

    # ---- optimizers: 6 consecutive steps on a fixed phi stream ------------
    n, d, steps = 12, 5, 6
    phis = rng.standard_normal((steps, n, d)) * np.array([1.0, 0.1, 3.0, 1e-3, 0.5, 2.0])[:, None, None]
    out = {"phis": phis}
    for tag, gd in (("adam", Adam(learning_rate=0.1, decay=0.999)),
                    ("adam_default", Adam()),
                    ("adagrad", Adagrad(learning_rate=0.05, decay=0.5, alpha=0.9))):
        upd = []
        for t in range(steps):
            upd.append(np.array(gd.update(phis[t].copy())))
        out[tag + "_updates"] = np.stack(upd)
        out[tag + "_final_lr"] = np.float64(gd.learning_rate)
        out[tag + "_n_iters"] = np.int64(gd.n_iters)
    np.savez(os.path.join(OUT, "optimizers.npz"), **out)

    # ---- converters: BNN-like variable set incl. the Variable_10 < Variable_2 sort
    names_shapes = [("model/Variable:0", []), ("model/Variable_1:0", []),
                    ("model/Variable_2:0", [3, 4]), ("model/Variable_3:0", [4]),
                    ("model/Variable_4:0", [4, 1]), ("model/Variable_5:0", []),
                    ("model/Variable_6:0", [2]), ("model/Variable_7:0", [2, 2]),
                    ("model/Variable_8:0", []), ("model/Variable_9:0", [1]),
                    ("model/Variable_10:0", [2, 3]), ("model/Variable_11:0", [])]
    vs = [FakeVar(nm, sh) for nm, sh in names_shapes]
    npart = 6
    dictionary = {v: rng.standard_normal([npart] + v._shape) for v in vs}
    array, access = conv.convert_dictionary_to_array(dictionary)
    back = conv.convert_array_to_dictionary(array, access)
    assert all(np.array_equal(back[v], dictionary[v]) for v in vs)
    np.savez(os.path.join(OUT, "converters.npz"),
             names=np.array([nm for nm, _ in names_shapes]),
             shapes=np.array([",".join(map(str, sh)) for _, sh in names_shapes]),
             array=array,
             starts=np.array([access[v][0] for v in vs]),
             stops=np.array([access[v][1] for v in vs]),
             **{"value_%d" % i: dictionary[v] for i, v in enumerate(vs)})

    # ---- linear-regression fixture + analytic posterior
    X = np.loadtxt(os.path.join(REF, "examples/linear_regression/data/data_X.csv"), delimiter=",")
    y = np.loadtxt(os.path.join(REF, "examples/linear_regression/data/data_y.csv"), delimiter=",")
    w = np.loadtxt(os.path.join(REF, "examples/linear_regression/data/data_w.csv"), delimiter=",")
    X2 = np.atleast_2d(X).T if X.ndim == 1 else X
    prec = X2.T @ X2 + np.eye(X2.shape[1])          # unit-variance likelihood, N(0,1) prior
    cov = np.linalg.inv(prec)
    mean = cov @ X2.T @ y
    np.savez(os.path.join(OUT, "linear_regression.npz"), X=X2.astype(np.float64), y=y,
             w_true=np.atleast_1d(w), post_mean=mean, post_cov=cov)
    print("posterior mean", mean, "var", np.diag(cov))


if __name__ == "__main__":
    main()
