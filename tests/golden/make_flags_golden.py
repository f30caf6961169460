"""Extract the command-line flags the reference's NeuralPoints / PointAggregator register (name, type, default, nargs) by parsing
its source with `ast` (the modules themselves are not importable here: PyCUDA, MinkowskiEngine), and store them as
tests/golden/reference_flags.json.  Run in the build container only (needs /root/reference):

    python tests/golden/make_flags_golden.py
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SGN_REFERENCE_ROOT", "/root/reference")
SOURCES = {"NeuralPoints": "models/neural_points/neural_points.py", "PointAggregator": "models/aggregators/point_aggregators.py"}


def flags_of(path, cls):
    out = []
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for m in node.body:
                if isinstance(m, ast.FunctionDef) and m.name == "modify_commandline_options":
                    for c in ast.walk(m):
                        if isinstance(c, ast.Call) and getattr(c.func, "attr", "") == "add_argument":
                            kw = {k.arg: k.value for k in c.keywords}
                            out.append({"flag": ast.literal_eval(c.args[0]), "type": ast.unparse(kw["type"]),
                                        "default": ast.literal_eval(kw["default"]),
                                        "nargs": ast.literal_eval(kw["nargs"]) if "nargs" in kw else None, "line": c.lineno})
    return sorted(out, key=lambda d: d["line"])


if __name__ == "__main__":
    data = {cls: flags_of(path, cls) for cls, path in SOURCES.items()}
    json.dump(data, open(os.path.join(HERE, "reference_flags.json"), "w"), indent=1)
    print({k: len(v) for k, v in data.items()})
