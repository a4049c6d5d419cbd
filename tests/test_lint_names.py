"""Static guard (no linters in the image): undefined names and local variables shadowing module-level imports anywhere in the product,
bench, tests and tools (tools/lint_names.py) -- the class of bug that only shows up when a GPU-only branch finally runs."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_undefined_or_shadowed_names():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "lint_names.py"), "prismer_b200", "bench.py", "__graft_entry__.py",
                        "tests", "oracle", "tools", "examples"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
