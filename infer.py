#!/usr/bin/env python
"""Whole-image (multi-scale) inference with the reference's interface
(/root/reference/infer.py:24-153): `predict_single_image`, `predict`,
`save_predictions`, `infer`, `main(data_dir, model_type='wesup', checkpoint=None,
output_dir=None, input_size=None, scales=(0.5,), num_workers=4, device=None)`.
The superpixel stage underneath is the CUDA path; there is no CPU mode.
"""
import warnings
from math import ceil
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

from wesup_b200 import cli
from wesup_b200.models import initialize_trainer
from wesup_b200.utils.data import SegmentationDataset

warnings.filterwarnings("ignore")


def predict_single_image(trainer, img, mask, output_size):
    """One forward at the given scale, nearest-resized back to `output_size` (:24-34)."""
    input_, target = trainer.preprocess(img, mask.long())
    with torch.no_grad():
        pred = trainer.model(input_)
    pred, _ = trainer.postprocess(pred, target)
    return F.interpolate(pred.float().unsqueeze(0), size=output_size, mode="nearest")


def _cross_opening(pred: np.ndarray, size: int = 9) -> np.ndarray:
    """Binary opening with the reference's cross-shaped structuring element (:84-92;
    note its `center = int((size+1)/2)` puts the cross one pixel off-centre)."""
    from scipy import ndimage
    selem = np.zeros((size, size))
    centre = int((size + 1) / 2)
    selem[centre, :] = 1
    selem[:, centre] = 1
    return ndimage.grey_opening(pred, footprint=selem.astype(bool))


def predict(trainer, dataset, input_size=None, scales=(0.5,), num_workers=4, device="cuda"):
    loader = torch.utils.data.DataLoader(dataset, num_workers=num_workers)
    scales = (scales,) if isinstance(scales, (int, float)) else tuple(scales)
    print(f"\nPredicting {len(dataset)} images with " + (f"input size {input_size}" if input_size else f"scales {scales}") + " ...")
    predictions = []
    for data in loader:
        img = data[0].to(device)
        mask = data[1].to(device).float()
        orig_size = (img.size(2), img.size(3))
        if input_size is not None:
            img = F.interpolate(img, size=input_size, mode="bilinear")
            mask = F.interpolate(mask, size=input_size, mode="nearest")
            prediction = predict_single_image(trainer, img, mask, orig_size)
        else:
            per_scale = []
            for scale in scales:
                size = [ceil(s * scale) for s in orig_size]
                img = F.interpolate(img, size=size, mode="bilinear")
                mask = F.interpolate(mask, size=size, mode="nearest")
                per_scale.append(predict_single_image(trainer, img, mask, orig_size))
            prediction = torch.cat(per_scale).mean(dim=0).round()
        prediction = prediction.squeeze().cpu().numpy()
        if input_size is None and len(scales) > 1:
            prediction = _cross_opening(prediction)
        predictions.append(prediction)
    return predictions


def save_predictions(predictions, dataset, output_dir="predictions"):
    print(f"\nSaving prediction to {output_dir} ...")
    output_dir = Path(output_dir)
    output_dir.mkdir(exist_ok=True)
    for pred, img_path in zip(predictions, dataset.img_paths):
        Image.fromarray(pred.astype("uint8") * 255).save(output_dir / f"{Path(img_path).stem}.png")


def infer(trainer, data_dir, output_dir=None, input_size=None, scales=(0.5,), num_workers=4, device="cuda"):
    trainer.model.eval()
    dataset = SegmentationDataset(data_dir, train=False)
    predictions = predict(trainer, dataset, input_size=input_size, scales=scales, num_workers=num_workers, device=device)
    if output_dir is not None:
        save_predictions(predictions, dataset, output_dir)
    return predictions


def main(data_dir, model_type="wesup", checkpoint=None, output_dir=None, input_size=None, scales=(0.5,),
         num_workers=4, device=None, **kwargs):
    if output_dir is None and checkpoint is not None:
        output_dir = Path(checkpoint).parent.parent / "results"
        output_dir.mkdir(exist_ok=True)
    device = device or "cuda"
    if checkpoint is not None:
        kwargs.setdefault("pretrained", False)       # every weight comes from the checkpoint: no ImageNet download
    trainer = initialize_trainer(model_type, device=device, **kwargs)
    if checkpoint is not None:
        trainer.load_checkpoint(checkpoint)
    return infer(trainer, data_dir, output_dir, input_size=input_size, scales=scales, num_workers=num_workers, device=device)


if __name__ == "__main__":
    cli.run(main)
