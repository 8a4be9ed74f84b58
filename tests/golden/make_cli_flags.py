"""Generates tests/golden/cli_flags.json and inference_flags.json: the flag tables of the reference CLIs
(/root/reference/train_textboost.py:49-428 ``parse_args`` and /root/reference/inference.py:21-43), extracted from
their ASTs in this container (the modules themselves cannot be imported: accelerate / diffusers / peft are not
installed).

    python tests/golden/make_cli_flags.py
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/train_textboost.py"


def extract(src):
    tree = ast.parse(open(src).read())
    flags = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            name = node.args[0].value
            spec = {}
            for k in node.keywords:
                if k.arg == "help":
                    continue
                try:
                    spec[k.arg] = ast.literal_eval(k.value)
                except ValueError:
                    spec[k.arg] = ast.unparse(k.value)  # type=str / int / float
            flags[name] = spec
    return flags


def main():
    flags = extract(SRC)
    with open(os.path.join(HERE, "cli_flags.json"), "w") as f:
        json.dump({"source": "train_textboost.py:49-428", "flags": flags}, f, indent=1, sort_keys=True)
    inf = extract("/root/reference/inference.py")
    with open(os.path.join(HERE, "inference_flags.json"), "w") as f:
        json.dump({"source": "inference.py:21-43", "flags": inf}, f, indent=1, sort_keys=True)
    print(len(flags), "+", len(inf), "flags")


if __name__ == "__main__":
    main()
