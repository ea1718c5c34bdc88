#!/usr/bin/env python3
"""Repository lint: cheap, dependency-free checks run in CI and by tests.

C++ / CUDA : no tabs, no trailing whitespace, lines <= 125 columns, include guards present
             in headers, no `using namespace` at header scope.
Python     : files compile; no tabs; lines <= 145 columns.
Parity: the reference wraps cpplint/pylint (tests/lint.py:18-60); neither is in this image.
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX_DIRS = ["include", "src", "apps", "cpp_tests", "pslite_b200/csrc"]
PY_DIRS = ["pslite_b200", "tracker", "tests", "scripts", "baseline"]
PY_FILES = ["bench.py", "__graft_entry__.py"]


def walk(dirs, exts):
    for d in dirs:
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            if "_ref" in base or "__pycache__" in base:
                continue
            for f in files:
                if f.endswith(exts):
                    yield os.path.join(base, f)


def lint_cxx(path, errs):
    is_header = path.endswith(".h")
    text = open(path, encoding="utf-8").read()
    if is_header and not re.search(r"#ifndef \w+_H_\n#define \w+_H_", text):
        errs.append(f"{path}: missing include guard")
    for i, line in enumerate(text.splitlines(), 1):
        if "\t" in line:
            errs.append(f"{path}:{i}: tab")
        if line != line.rstrip():
            errs.append(f"{path}:{i}: trailing whitespace")
        if len(line) > 125:
            errs.append(f"{path}:{i}: line longer than 125 columns ({len(line)})")
        if is_header and re.match(r"using namespace \w+;", line):
            errs.append(f"{path}:{i}: using-directive at header scope")


def lint_py(path, errs):
    src = open(path, encoding="utf-8").read()
    try:
        compile(src, path, "exec")
    except SyntaxError as e:
        errs.append(f"{path}:{e.lineno}: {e.msg}")
    for i, line in enumerate(src.splitlines(), 1):
        if "\t" in line:
            errs.append(f"{path}:{i}: tab")
        if len(line) > 145:
            errs.append(f"{path}:{i}: line longer than 145 columns ({len(line)})")


def main() -> int:
    errs: list[str] = []
    for p in walk(CXX_DIRS, (".h", ".cc", ".cu")):
        lint_cxx(p, errs)
    for p in list(walk(PY_DIRS, (".py",))) + [os.path.join(ROOT, f) for f in PY_FILES]:
        lint_py(p, errs)
    for e in errs[:200]:
        print(e)
    print(f"lint: {len(errs)} problem(s)")
    return 1 if errs else 0


if __name__ == "__main__":
    sys.exit(main())
