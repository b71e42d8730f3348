"""Record the public signatures of the reference front end (nvdiffrast/torch/ops.py) as a small fixture, so that the
drop-in surface can be checked where /root/reference is not available.  Run in the build container:

    python tests/golden/make_api_signatures.py        # writes tests/golden/api_signatures.json
"""
import ast
import json
import os

SRC = "/root/reference/nvdiffrast/torch/ops.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_signatures.json")


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return [{"name": n, "default": d} for n, d in zip(names, defaults)]


def main():
    tree = ast.parse(open(SRC).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and not node.name.startswith("_"):
            out[node.name] = sig(node)
        if isinstance(node, ast.ClassDef) and not node.name.startswith("_"):
            methods = {m.name: sig(m) for m in node.body if isinstance(m, ast.FunctionDef) and (not m.name.startswith("_") or m.name in ("__init__", "__enter__", "__exit__"))}
            out[node.name] = {"methods": methods}
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, sorted(out))


if __name__ == "__main__":
    main()
