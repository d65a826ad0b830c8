#!/usr/bin/env python
"""Undefined-name check for code paths that cannot run on this box (GPU-only branches, gated tests): no linter is installed
in the image, so this walks every function with `symtable` and reports global-scope names that are neither module-level
bindings, imports, nor builtins.

    python tools/lint_names.py            # package, tests, tools, bench.py, __graft_entry__.py
    python tools/lint_names.py path ...   # explicit files / directories
"""
from __future__ import annotations

import ast
import builtins
import os
import symtable
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def module_bindings(tree: ast.Module) -> set[str]:
    names: set[str] = set()
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                if a.name == "*":
                    names.add("*")
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, ast.Global):
            names.update(node.names)
    for node in tree.body:
        for sub in ast.walk(node) if not isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)) else [node]:
            if isinstance(sub, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                names.add(sub.name)
            elif isinstance(sub, ast.Name) and isinstance(sub.ctx, (ast.Store, ast.Del)):
                names.add(sub.id)
            elif isinstance(sub, ast.ExceptHandler) and sub.name:
                names.add(sub.name)
    return names


def check_file(path: str) -> list[str]:
    src = open(path, encoding="utf-8").read()
    try:
        tree = ast.parse(src, path)
        table = symtable.symtable(src, path, "exec")
    except SyntaxError as e:
        return [f"{path}:{e.lineno}: syntax error: {e.msg}"]
    known = module_bindings(tree) | set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__path__", "__spec__", "__builtins__"}
    if "*" in known:
        return []
    problems = []
    # line numbers of every Name load, for the report
    first_use: dict[str, int] = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load):
            first_use.setdefault(node.id, node.lineno)

    def visit(t: symtable.SymbolTable):
        for sym in t.get_symbols():
            if t.get_type() == "module":
                if sym.is_referenced() and not sym.is_assigned() and not sym.is_imported() and sym.get_name() not in known:
                    problems.append(f"{path}:{first_use.get(sym.get_name(), 0)}: undefined name {sym.get_name()!r}")
            elif sym.is_global() and sym.is_referenced() and not sym.is_assigned() and sym.get_name() not in known:
                problems.append(f"{path}:{first_use.get(sym.get_name(), 0)}: undefined name {sym.get_name()!r} in {t.get_name()}()")
        for child in t.get_children():
            visit(child)
    visit(table)
    return sorted(set(problems))


def collect(paths: list[str]) -> list[str]:
    out = []
    for p in paths:
        if os.path.isdir(p):
            for d, dirs, files in os.walk(p):
                dirs[:] = [x for x in dirs if x not in ("__pycache__", "_ref", "build", ".git", "gpurun_out")]
                out += [os.path.join(d, f) for f in files if f.endswith(".py")]
        elif p.endswith(".py"):
            out.append(p)
    return sorted(out)


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    targets = argv or [os.path.join(ROOT, x) for x in ("deeplearning_cfn_b200", "tests", "tools", "bench.py", "__graft_entry__.py")]
    problems = []
    files = collect(targets)
    for f in files:
        problems += check_file(f)
    for p in problems:
        print(p)
    print(f"{len(files)} files checked, {len(problems)} undefined names")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
