"""NumPy float64 interpreter of the captured-model expression DAG (test helper):
lets the CPU suite check gen/autodiff.py without a GPU."""
import math

import numpy as np
from scipy.special import gammaln


def evaluate(e, env, cache=None):
    """env: site index -> value, ('arg', i) -> value."""
    cache = {} if cache is None else cache
    if e._id in cache:
        return cache[e._id]
    op = e.op
    ins = [evaluate(i, env, cache) for i in e.ins]
    if op == "const":
        v = float(e.attr) if e.dtype == "f32" else int(e.attr)
    elif op == "constvec":
        v = np.array(e.attr, dtype=np.float64)
    elif op == "site":
        v = np.asarray(env[e.attr], dtype=np.float64)
    elif op == "arg":
        v = np.asarray(env[("arg", e.attr["index"])], dtype=np.float64)
    elif op == "chain_step":
        v = env["step_size"]
    elif op == "row":
        v = ins[0][int(ins[1])]
    elif op == "gather1":
        v = ins[0][int(ins[1])]
    elif op == "elem":
        v = ins[0][int(e.attr)]
    elif op == "sum":
        v = np.sum(ins[0])
    elif op == "cast":
        v = np.asarray(ins[0], dtype=np.float64) if e.dtype == "f32" else np.asarray(ins[0]).astype(np.int64)
    elif op == "where":
        v = np.where(ins[0], ins[1], ins[2])
    else:
        f = {
            "add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "pow": np.power,
            "min": np.minimum, "max": np.maximum, "lt": np.less, "le": np.less_equal, "gt": np.greater,
            "ge": np.greater_equal, "eq": np.equal, "ne": np.not_equal, "and": np.logical_and, "or": np.logical_or,
            "neg": np.negative, "exp": np.exp, "log": np.log, "sqrt": np.sqrt, "abs": np.abs, "tanh": np.tanh,
            "sigmoid": lambda x: 1 / (1 + np.exp(-x)), "log1p": np.log1p, "expm1": np.expm1, "square": np.square,
            "floor": np.floor, "sin": np.sin, "cos": np.cos, "softplus": lambda x: np.logaddexp(0, x),
            "lgamma": gammaln, "reciprocal": lambda x: 1 / x, "logical_not": np.logical_not,
        }[op]
        with np.errstate(all="ignore"):
            v = f(*ins)
    cache[e._id] = v
    return v
