"""Poor man's lint for the GPU-only Python (bench.py, lidog_b200/): every name a function loads must be bound in that
function (or an enclosing one), at module level, or be a builtin.  The product path cannot run in the CPU test
environment, so a typo there would otherwise surface only on the GPU box."""
import ast
import builtins
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bound_names(node):
    names = set()
    for m in ast.walk(node):
        if isinstance(m, ast.Name) and isinstance(m.ctx, (ast.Store, ast.Del)):
            names.add(m.id)
        elif isinstance(m, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(m.name)
        elif isinstance(m, (ast.Import, ast.ImportFrom)):
            for a in m.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(m, ast.ExceptHandler) and m.name:
            names.add(m.name)
        elif isinstance(m, ast.arg):
            names.add(m.arg)
        elif isinstance(m, (ast.Global, ast.Nonlocal)):
            names.update(m.names)
    return names


def _undefined(path):
    tree = ast.parse(open(path).read(), path)
    module_names = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for n in tree.body:  # module level: everything bound outside function bodies
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            module_names.add(n.name)
        else:
            module_names |= _bound_names(n)
    bad = []
    for fn in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef))]:
        local = _bound_names(fn)  # includes nested functions' names: a superset, which is fine for this check
        for m in ast.walk(fn):
            if isinstance(m, ast.Name) and isinstance(m.ctx, ast.Load) and m.id not in local and m.id not in module_names:
                bad.append((os.path.relpath(path, ROOT), m.lineno, m.id))
    # class bodies may bind names used by their methods' defaults etc.; enclosing-function names are handled by the
    # superset above only for the innermost function, so re-check against all enclosing functions' bindings
    enclosing = {}
    for outer in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef))]:
        for inner in ast.walk(outer):
            if inner is not outer and isinstance(inner, (ast.FunctionDef, ast.AsyncFunctionDef)):
                enclosing.setdefault(id(inner), set()).update(_bound_names(outer))
    out = []
    for rel, line, name in bad:
        ok = False
        for fn in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef))]:
            if fn.lineno <= line <= max(getattr(fn, "end_lineno", fn.lineno), fn.lineno) and name in enclosing.get(id(fn), ()):
                ok = True
                break
        if not ok:
            out.append((rel, line, name))
    return sorted(set(out))


def test_gpu_only_python_has_no_unbound_names():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += glob.glob(os.path.join(ROOT, "lidog_b200", "**", "*.py"), recursive=True)
    files += glob.glob(os.path.join(ROOT, "tools", "*.py"))
    problems = []
    for f in sorted(files):
        problems += _undefined(f)
    assert not problems, problems
