"""Record-directory bookkeeping (/root/reference/utils/record.py:19-52).  The
reference also snapshots its sources and plots curves with matplotlib; neither
touches the hot path and matplotlib is not a dependency here."""
from __future__ import annotations

import json
import os
from datetime import datetime
from pathlib import Path


def prepare_record_dir():
    root = Path(os.environ["RECORD_ROOT"]).expanduser() if os.environ.get("RECORD_ROOT") else Path.home() / "records"
    record_dir = root / datetime.now().strftime("%Y%m%d-%I%M-%p")
    (record_dir / "checkpoints").mkdir(parents=True, exist_ok=True)
    return record_dir


def save_params(record_dir, params):
    params_dir = Path(record_dir) / "params"
    params_dir.mkdir(exist_ok=True)
    with open(params_dir / f"{len(list(params_dir.iterdir()))}.json", "w") as fp:
        json.dump(params, fp, indent=4)
