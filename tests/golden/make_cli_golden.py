"""Generates tests/golden/cli_flags.json: every `add_argument` of the reference's two stage CLIs
(/root/reference/main_img_denoising.py:152-217, /root/reference/main_denoiser.py:25-78) -- flag names with their `type`,
`default`, `action`, `nargs`, `choices` expressions -- read from the source with `ast` (the modules themselves import timm /
tinycudann and cannot be executed here).  tests/test_store_cpu.py::test_cli_flags_match_the_reference checks that the
drop-in CLIs accept the same flags with the same defaults.  Also tests/golden/api_signatures.json: the public method
signatures of the four `dvt/models` classes on the hot paths (reference dvt/models/{vit_wrapper,neural_feature_field,
offline_denoiser,online_denoiser}.py) and MODEL_LIST, checked by test_model_api_signatures_match_the_reference.

Run in the build container (needs /root/reference):  python tests/golden/make_cli_golden.py"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("type", "default", "action", "nargs", "choices")


def flags(path):
    out = {}
    for n in ast.walk(ast.parse(open(path).read())):
        if isinstance(n, ast.Call) and getattr(n.func, "attr", "") == "add_argument":
            names = [a.value for a in n.args if isinstance(a, ast.Constant) and isinstance(a.value, str)]
            kw = {k.arg: ast.unparse(k.value) for k in n.keywords if k.arg in KEYS}
            for nm in names:
                out[nm] = kw
    return out


MODEL_FILES = ("vit_wrapper.py", "neural_feature_field.py", "offline_denoiser.py", "online_denoiser.py")


def signatures(path):
    """Public methods (and __init__) of every top-level class: positional parameter names and default expressions; plus
    the literal MODEL_LIST where the file defines one."""
    out = {}
    for c in ast.parse(open(path).read()).body:
        if isinstance(c, ast.ClassDef):
            for f in c.body:
                if isinstance(f, ast.FunctionDef) and (not f.name.startswith("_") or f.name == "__init__"):
                    out[f"{c.name}.{f.name}"] = {"args": [a.arg for a in f.args.args],
                                                 "defaults": [ast.unparse(d) for d in f.args.defaults]}
        if isinstance(c, ast.Assign) and any(getattr(t, "id", "") == "MODEL_LIST" for t in c.targets):
            out["MODEL_LIST"] = ast.literal_eval(c.value)
    return out


if __name__ == "__main__":
    gold = {f: flags(os.path.join("/root/reference", f)) for f in ("main_img_denoising.py", "main_denoiser.py")}
    with open(os.path.join(HERE, "cli_flags.json"), "w") as fh:
        json.dump(gold, fh, indent=1, sort_keys=True)
    print({k: len(v) for k, v in gold.items()})
    api = {f: signatures(os.path.join("/root/reference/dvt/models", f)) for f in MODEL_FILES}
    with open(os.path.join(HERE, "api_signatures.json"), "w") as fh:
        json.dump(api, fh, indent=1, sort_keys=True)
    print({k: len(v) for k, v in api.items()})
