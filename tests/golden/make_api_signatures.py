"""Extracts the call signatures of the reference's public API (what `jpc/__init__.py` re-exports) by parsing its sources
with `ast` -- nothing is imported (JAX is not installable here) and nothing is copied except parameter names and
literal defaults.  Output: tests/golden/api_signatures.json, checked by tests/test_api_surface.py.
Run in the build container:  python tests/golden/make_api_signatures.py /root/reference"""
import ast
import warnings
import json
import os
import sys


def literal(node):
    if node is None:
        return "<required>"
    try:
        return repr(ast.literal_eval(node))
    except Exception:
        return "<expr>"   # e.g. Heun(), PIDController(rtol=1e-3, atol=1e-3)


def main(root):
    warnings.simplefilter("ignore", SyntaxWarning)  # (LaTeX backslashes in the docstrings being parsed)
    init = ast.parse(open(os.path.join(root, "jpc", "__init__.py")).read())
    exported = set()
    for node in ast.walk(init):
        if isinstance(node, ast.ImportFrom):
            for a in node.names:
                exported.add(a.asname or a.name)
    sigs = {}
    for dirpath, _, files in os.walk(os.path.join(root, "jpc")):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in tree.body:
                if isinstance(node, ast.FunctionDef) and node.name in exported:
                    a = node.args
                    pos = [x.arg for x in a.posonlyargs + a.args]
                    pos_defaults = [None] * (len(pos) - len(a.defaults)) + list(a.defaults)
                    params = [[n, "positional", literal(d) if d is not None else "<required>"] for n, d in zip(pos, pos_defaults)]
                    params += [[x.arg, "keyword", literal(d)] for x, d in zip(a.kwonlyargs, a.kw_defaults)]
                    sigs[node.name] = {"file": os.path.relpath(os.path.join(dirpath, f), root) + f":{node.lineno}", "params": params}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_signatures.json")
    json.dump({"exported": sorted(exported), "signatures": sigs}, open(out, "w"), indent=1, sort_keys=True)
    print(len(sigs), "signatures ->", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
