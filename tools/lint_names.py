"""Tiny undefined-name check (no pyflakes in the image): every Name that is loaded must be
bound somewhere in an enclosing function / module scope, be a builtin, or a comprehension
variable.  Catches the typos that would otherwise cost a GPU call to find.

  python tools/lint_names.py nerfies_b200 bench.py tools tests __graft_entry__.py
"""
import ast, builtins, os, sys


def bound_names(node):
  names = set()
  for n in ast.walk(node):
    if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
      names.add(n.id)
    elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
      names.add(n.name)
    elif isinstance(n, (ast.Import, ast.ImportFrom)):
      for a in n.names:
        names.add((a.asname or a.name).split('.')[0])
    elif isinstance(n, ast.arg):
      names.add(n.arg)
    elif isinstance(n, ast.ExceptHandler) and n.name:
      names.add(n.name)
    elif isinstance(n, (ast.Global, ast.Nonlocal)):
      names.update(n.names)
  return names


def check(path):
  tree = ast.parse(open(path).read(), path)
  ok = set(dir(builtins)) | bound_names(tree) | {'__file__', '__name__', '__doc__'}
  bad = []
  for n in ast.walk(tree):
    if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in ok:
      bad.append((n.lineno, n.id))
  return bad


def main(args):
  files = []
  for a in args:
    if os.path.isdir(a):
      for root, _, fs in os.walk(a):
        files += [os.path.join(root, f) for f in fs if f.endswith('.py')]
    else:
      files.append(a)
  n = 0
  for f in sorted(files):
    for line, name in check(f):
      print(f'{f}:{line}: undefined name {name!r}')
      n += 1
  return 1 if n else 0


if __name__ == '__main__':
  sys.exit(main(sys.argv[1:]))
