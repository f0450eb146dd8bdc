"""Host-side helpers mirrored from the reference's `utils` package surface
(/root/reference/utils/__init__.py:4-19): the 0-dim "no tensor" sentinel and
its test are part of the data contract (utils/data.py:22,144)."""
import torch


def underline(content, style="-"):
    return f"{content}\n{style * len(content.strip())}"


def empty_tensor():
    """The reference's "no mask" sentinel: a 0-dim tensor."""
    return torch.tensor(0)


def is_empty_tensor(t):
    return t.dim() == 0
