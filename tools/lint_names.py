"""Tiny static check (no linters in this image): per function, (1) names that are read but are neither local, enclosing, module-level
nor builtins (-> NameError at run time) and (2) local variables that shadow a module-level import (-> UnboundLocalError when the import
is used before the assignment, the bug class caught in prismer_caption.py at the end of round 1).

    python tools/lint_names.py prismer_b200 bench.py __graft_entry__.py tests oracle tools examples
"""
import ast
import builtins
import os
import sys


def check(path):
    src = open(path).read()
    tree = ast.parse(src)
    mod_names, imports = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}, set()
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                nm = (a.asname or a.name).split(".")[0]
                mod_names.add(nm)
                if n in tree.body or any(n in getattr(b, "body", []) for b in tree.body if isinstance(b, (ast.If, ast.Try))):
                    imports.add(nm)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            mod_names.add(n.name)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            mod_names.add(n.id)        # over-approximation: any assigned name anywhere counts as "known"
        elif isinstance(n, ast.arg):
            mod_names.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            mod_names.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            mod_names.update(n.names)
    out = []
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in mod_names:
            out.append(f"{path}:{n.lineno}: undefined name '{n.id}'")
    for fn in ast.walk(tree):
        if not isinstance(fn, (ast.FunctionDef, ast.AsyncFunctionDef)):
            continue
        assigned = set()
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
                assigned.add(n.id)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):
                for a in n.names:
                    assigned.discard((a.asname or a.name).split(".")[0])   # a local import re-binds on purpose
        local_imports = {(a.asname or a.name).split(".")[0] for n in ast.walk(fn) if isinstance(n, (ast.Import, ast.ImportFrom)) for a in n.names}
        for nm in sorted((assigned & imports) - local_imports):
            out.append(f"{path}:{fn.lineno}: local variable '{nm}' in {fn.name}() shadows the module-level import")
    return out


def main():
    problems = []
    for target in sys.argv[1:]:
        files = [target] if target.endswith(".py") else [os.path.join(r, f) for r, _, fs in os.walk(target) for f in fs if f.endswith(".py")]
        for f in sorted(files):
            problems += check(f)
    print("\n".join(problems) if problems else "no problems found")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
