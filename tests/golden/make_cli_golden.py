"""Generates tests/golden/cli_flags.json: every `add_argument` of the reference's two stage CLIs
(/root/reference/main_img_denoising.py:152-217, /root/reference/main_denoiser.py:25-78) -- flag names with their `type`,
`default`, `action`, `nargs`, `choices` expressions -- read from the source with `ast` (the modules themselves import timm /
tinycudann and cannot be executed here).  tests/test_store_cpu.py::test_cli_flags_match_the_reference checks that the
drop-in CLIs accept the same flags with the same defaults.

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


if __name__ == "__main__":
    gold = {f: flags(os.path.join("/root/reference", f)) for f in ("main_img_denoising.py", "main_denoiser.py")}
    with open(os.path.join(HERE, "cli_flags.json"), "w") as fh:
        json.dump(gold, fh, indent=1, sort_keys=True)
    print({k: len(v) for k, v in gold.items()})
