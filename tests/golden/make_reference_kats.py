#!/usr/bin/env python
"""Extract the known-answer vectors held by the reference's own tests.

TensorFlow 1.x cannot run in this container, so the reference tests cannot be
*executed*; but their inputs and expected outputs are plain literals.  This
script parses the reference test sources with ``ast`` (nothing is imported
from them), evaluates the numpy-only statements of each test method with a
recording stand-in for ``self``, and writes every (inputs, expected) tuple to
``reference_kats.json`` next to this file.  Run it in the build container
(where /root/reference exists); the JSON is committed and is what the tests
read -- /root/reference is absent on the GPU box.

Sources (relative to /root/reference/src/e2eflow/test/):
  ops/correlation.py:30-69, ops/backward_warp.py:27-101, ops/downsample.py:8-22,
  test_image_warp.py:23-97, test_losses.py:11-68,97-121.
"""
import ast
import json
import os
import sys

import numpy as np

REF = os.environ.get("UNFLOW_REFERENCE", "/root/reference")
TEST_DIR = os.path.join(REF, "src", "e2eflow", "test")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def _tolist(v):
    if isinstance(v, np.ndarray):
        return {"shape": list(v.shape), "data": v.astype(np.float64).ravel().tolist()}
    if isinstance(v, (list, tuple)):
        a = np.asarray(v, dtype=np.float64)
        return {"shape": list(a.shape), "data": a.ravel().tolist()}
    if isinstance(v, (int, float, np.floating, np.integer)):
        return float(v)
    if v is None:
        return None
    raise TypeError(type(v))


class _Recorder:
    def __init__(self):
        self.calls = []


def run_method(func_node, source_name):
    """Execute the numpy-only statements of a test method; record self.<m>(...) calls."""
    ns = {"np": np}
    rec = _Recorder()
    for stmt in func_node.body:
        if isinstance(stmt, ast.Assign):
            try:
                exec(compile(ast.Module([stmt], []), source_name, "exec"), ns)
            except Exception:
                pass  # depends on TF objects
        elif isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Call):
            call = stmt.value
            f = call.func
            if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "self":
                args = []
                for a in call.args:
                    try:
                        args.append(_tolist(eval(compile(ast.Expression(a), source_name, "eval"), ns)))
                    except Exception:
                        args.append({"expr": ast.unparse(a)})
                kwargs = {}
                for kw in call.keywords:
                    kwargs[kw.arg] = eval(compile(ast.Expression(kw.value), source_name, "eval"), ns)
                rec.calls.append({"method": f.attr, "args": args, "kwargs": kwargs,
                                  "line": stmt.lineno})
    return rec.calls, ns


def extract(relpath, wanted):
    path = os.path.join(TEST_DIR, relpath)
    with open(path) as fh:
        tree = ast.parse(fh.read(), path)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            calls, ns = run_method(node, path)
            entry = {"source": "src/e2eflow/test/%s:%d" % (relpath, node.lineno), "calls": calls}
            extra = wanted[node.name]
            if extra:
                entry["vars"] = {k: _tolist(ns[k]) for k in extra}
            out[node.name] = entry
    missing = set(wanted) - set(out)
    if missing:
        raise SystemExit("missing tests in %s: %s" % (relpath, sorted(missing)))
    return out


def main():
    kats = {
        "correlation": extract("ops/correlation.py",
                               {"test_correlation_trivial": None, "test_correlation_batch": None}),
        "backward_warp": extract("ops/backward_warp.py",
                                 {"test_move": None, "test_batches": None, "test_interpolate": None}),
        "image_warp": extract("test_image_warp.py",
                              {"test_move": None, "test_batches": None, "test_interpolate": None}),
        "downsample": extract("ops/downsample.py", {"test_downsample": ["first", "second"]}),
        "losses": extract("test_losses.py",
                          {"test_smoothness_deltas": ["flow"],
                           "test_create_outgoing_mask_all_directions": ["flow"],
                           "test_create_outgoing_mask_large_movement": ["flow"],
                           "test_gradient_loss": ["im1", "im2", "mask"]}),
    }
    with open(OUT, "w") as fh:
        json.dump(kats, fh, indent=1, sort_keys=True)
    n = sum(len(v) for v in kats.values())
    print("wrote %s (%d reference tests)" % (OUT, n))


if __name__ == "__main__":
    sys.exit(main())
