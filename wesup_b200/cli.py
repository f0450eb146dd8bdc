"""Tiny python-fire stand-in for the entry points (the reference's CLIs are
`fire.Fire(fn)`, e.g. /root/reference/train.py:31; `fire` is not a dependency
here): positional words fill the function's positional parameters, every
`--key value` / `--key=value` / bare `--flag` becomes a keyword argument, and
values go through `ast.literal_eval` when they parse (so `--scales "(0.5,1)"`
and `--epochs 3` behave as they do under fire)."""
from __future__ import annotations

import ast
import sys
from typing import Callable, Sequence


def _value(text: str):
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


def parse(argv: Sequence[str]):
    args, kwargs = [], {}
    i = 0
    while i < len(argv):
        word = argv[i]
        if word.startswith("--"):
            key = word[2:]
            if "=" in key:
                key, val = key.split("=", 1)
                kwargs[key.replace("-", "_")] = _value(val)
            elif i + 1 < len(argv) and not argv[i + 1].startswith("--"):
                kwargs[key.replace("-", "_")] = _value(argv[i + 1])
                i += 1
            else:
                kwargs[key.replace("-", "_")] = True
        else:
            args.append(_value(word))
        i += 1
    return args, kwargs


def run(fn: Callable, argv: Sequence[str] | None = None):
    args, kwargs = parse(sys.argv[1:] if argv is None else argv)
    return fn(*args, **kwargs)
