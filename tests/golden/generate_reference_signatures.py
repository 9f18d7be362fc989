"""Writes tests/golden/reference_signatures.json: argument names and defaults of every public function the reference
package exports (parsed from /root/reference with ast -- nothing is imported or copied).  Run in the build container."""
import ast
import glob
import json
import os

REF = "/root/reference/mlx_audio_primitives"


def signatures(paths):
    out = {}
    for p in paths:
        for n in ast.parse(open(p).read()).body:
            if isinstance(n, ast.FunctionDef) and not n.name.startswith("_"):
                a = n.args
                names = [x.arg for x in a.posonlyargs + a.args]
                defs = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
                kw = [[x.arg, ast.unparse(d) if d else None] for x, d in zip(a.kwonlyargs, a.kw_defaults)]
                out[n.name] = {"args": [list(z) for z in zip(names, defs)], "kwonly": kw}
    return out


if __name__ == "__main__":
    exported = []
    for n in ast.walk(ast.parse(open(os.path.join(REF, "__init__.py")).read())):
        if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "__all__":
            exported = [e.value for e in n.value.elts]
    sig = signatures(glob.glob(os.path.join(REF, "*.py")))
    out = {k: sig[k] for k in exported if k in sig}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_signatures.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "signatures")
