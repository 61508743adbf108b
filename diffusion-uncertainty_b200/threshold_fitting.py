"""Offline per-timestep / per-pixel threshold fitting with the reference's on-disk contract — drop-in for the
body of scripts/compute_threshold_pixel_wise.py:66-116 (`main`) and :118-164 (`load_uncertainty_datasets`).

Input: dataset folders written by the uncertainty-map generation (`uncertainty_*.pth` `[N, T_uc, C, H, W]`, `gen_images_*.pth`,
`args.yaml` with at least `dataset` and `scheduler_type`).  Output, as the reference writes it and as
scripts/generate_images_with_uncertainty_threshold.py:157-176 reads it back:
    <results>/thresholds/<dataset>/thresholds_<scheduler_type>_perc=<perc>.pth     fp16 `[T_uc, C, H, W]` (maps are cast `.half()`, :143)
    <results>/thresholds/<dataset>/config_<scheduler_type>_perc=<perc>.yaml        the arguments + `dataset_config` + `dataset_folders`
The select itself — `argsort(dim=0)[int(N * perc)]` + `gather` per timestep (:90-100, 147-157) — is du_column_kth on the device
(ops.fit_pixel_thresholds); files with fewer than 100 samples are skipped (:145-146); several files / folders are merged by the
same select over their per-file thresholds (:77-100).
"""
from __future__ import annotations

import glob
import os
from typing import Dict, List, Optional, Sequence

import torch
import yaml

from . import ops


def load_uncertainty_datasets(dataset_folders: Sequence[str], perc: float, device="cuda") -> List[torch.Tensor]:
    """Per uncertainty file: the `[1, T_uc, C, H, W]` thresholds (fp16, on `device`) — reference :118-164."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"device {device}: the uncertainty path has no CPU fallback")
    out = []
    for folder in dataset_folders:
        folder = str(folder)
        files = sorted(glob.glob(os.path.join(folder, "uncertainty_*.pth")))
        assert len(files) > 0, f"no uncertainty files found in {folder}"
        assert len(glob.glob(os.path.join(folder, "gen_images_*.pth"))) > 0, f"no gen_images files found in {folder}"
        for path in files:
            u = torch.load(path, map_location="cpu").half()
            if u.shape[0] < 100:
                continue
            out.append(ops.fit_pixel_thresholds(u.to(device), perc).unsqueeze(0))
    return out


def fit_thresholds(dataset_folders: Sequence[str], perc: float, device="cuda") -> torch.Tensor:
    """`timestep_thresholds` of the reference's main (:75-100)."""
    parts = load_uncertainty_datasets(dataset_folders, perc, device)
    if len(parts) == 1:
        return parts[0].squeeze(0)
    return ops.fit_pixel_thresholds(torch.cat(parts, dim=0), perc)


def fit_and_save_thresholds(dataset_folders: Sequence[str], perc: float, results_dir: str, device="cuda",
                            extra_args: Optional[Dict] = None) -> Dict[str, str]:
    """Fit and write the two files; returns their paths.  `results_dir` is the reference's RESULTS directory."""
    with open(os.path.join(str(dataset_folders[0]), "args.yaml")) as f:
        config = yaml.safe_load(f)
    thr = fit_thresholds(dataset_folders, perc, device).cpu()
    folder = os.path.join(str(results_dir), "thresholds", str(config["dataset"]))
    os.makedirs(folder, exist_ok=True)
    stem = f'{config["scheduler_type"]}_perc={perc}'
    thr_path = os.path.join(folder, f"thresholds_{stem}.pth")
    torch.save(thr, thr_path)
    args_dict = dict(extra_args or {})
    args_dict.update({"perc": perc, "dataset_config": config, "dataset_folders": [str(x) for x in dataset_folders]})
    cfg_path = os.path.join(folder, f"config_{stem}.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(args_dict, f)
    return {"thresholds": thr_path, "config": cfg_path}
