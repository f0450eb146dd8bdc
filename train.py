#!/usr/bin/env python
"""Training entry point with the reference's interface (/root/reference/train.py:14-32):

    python train.py <dataset_path> [--model wesup] [--epochs E] [--smoke] [--any_config_key value ...]

Every `--key value` becomes a keyword argument of `initialize_trainer` and
`trainer.train`, as under python-fire.  Additive: launched with torchrun
(`python -m torch.distributed.run --nproc-per-node N train.py ...`) it trains
data-parallel, one process per GPU, gradients averaged with one NCCL all-reduce
per iteration (wesup_b200.parallel).
"""
import logging
from shutil import rmtree

from wesup_b200 import cli, parallel
from wesup_b200.models import initialize_trainer
from wesup_b200.utils.metrics import accuracy, dice


def fit(dataset_path, model="wesup", **kwargs):
    logger = logging.getLogger("Train")
    logger.setLevel(logging.DEBUG)
    if not logger.handlers:
        logger.addHandler(logging.StreamHandler())
    rank, world, local = parallel.init_from_env()
    if world > 1:
        kwargs.setdefault("device", f"cuda:{local}")
    trainer = initialize_trainer(model, logger=logger, **kwargs)
    if world > 1:
        trainer.enable_data_parallel()
    try:
        trainer.train(dataset_path, metrics=[accuracy, dice], **kwargs)
    finally:
        if kwargs.get("smoke") and trainer.record_dir is not None:
            rmtree(trainer.record_dir, ignore_errors=True)
    return trainer


if __name__ == "__main__":
    cli.run(fit)
