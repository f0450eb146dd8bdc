"""`models` package surface of the reference (/root/reference/models/__init__.py:9-19)."""
from .wesup import WESUP, WESUPConfig, WESUPPixelInference, WESUPTrainer


def initialize_trainer(model_type, **kwargs):
    """Factory used by train.py / infer.py / infer_tile.py: only 'wesup' exists."""
    if model_type != "wesup":
        raise ValueError(f"Unsupported model: {model_type}")
    kwargs = {**WESUPConfig().to_dict(), **kwargs}
    return WESUPTrainer(WESUP(**kwargs), **kwargs)
