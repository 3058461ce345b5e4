"""Repo hygiene the scope contract asks for: the product package never imports the oracle, and the
required top-level files exist."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyro_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
                # the test-only host harness is never loaded by product code
                if f.endswith(".py") and f != "_build.py" and "hostcheck" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_required_files():
    for rel in ("include/pyro_b200.h", "bench.py", "__graft_entry__.py", "DESIGN.md",
                "INTEGRATION.md", "oracle/__init__.py", "tests/golden/make_golden.py"):
        assert os.path.exists(os.path.join(ROOT, rel)), rel
