"""The product path may never import, call or link the CPU oracle (it is test infrastructure),
and must have no CPU fallback for the ops."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product_files():
    for base, _, files in os.walk(os.path.join(ROOT, "unflow_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                yield os.path.join(base, f)
    yield os.path.join(ROOT, "include", "unflow.h")


def test_product_never_references_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|oracle_ops|sys\.path.*oracle", re.M)
    for path in _product_files():
        src = open(path).read()
        # citations of the reference in comments are fine; imports / links of the oracle are not
        assert not pat.search(src), "%s references the oracle" % path


def test_bench_and_smoke_do_not_read_the_reference_tree():
    for f in ("bench.py", "__graft_entry__.py"):
        src = open(os.path.join(ROOT, f)).read()
        assert "/root/reference" not in src, f
