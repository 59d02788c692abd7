"""Configuration dataclasses with the reference's field names.

``DatasetConfig`` mirrors ``datasets/config.py:6-14`` (class ``Dataset``) and ``Config`` mirrors
``depth_pose_prediction/config.py:6-32`` (class ``DepthPosePrediction``).  The predictor only
duck-types its two config arguments, so the reference's own dataclass instances (as produced by
``config/config_parser.py``) can be passed unchanged.
"""
from __future__ import annotations

import dataclasses
from pathlib import Path
from typing import Optional, Tuple, Union


@dataclasses.dataclass
class DatasetConfig:
    dataset: str = 'Kitti'
    config_file: Path = Path('config.yaml')
    dataset_path: Optional[Path] = None
    scales: Optional[Tuple[int, ...]] = (0, 1, 2, 3)
    height: Optional[int] = 192
    width: Optional[int] = 640
    frame_ids: Tuple[int, ...] = (0, -1, 1)


@dataclasses.dataclass
class Config:
    config_file: Path = Path('config.yaml')
    train_set: Optional[Union[Tuple[int, ...], int, str]] = (0,)
    val_set: Optional[Union[Tuple[int, ...], Tuple[str, ...], int, str]] = (9,)
    resnet_depth: int = 18
    resnet_pose: int = 18
    resnet_pretrained: bool = False
    scales: Tuple[int, ...] = (0, 1, 2, 3)
    learning_rate: float = 1e-4
    scheduler_step_size: int = 15
    batch_size: int = 3
    num_workers: int = 0
    num_epochs: int = 20
    min_depth: Optional[float] = 0.1
    max_depth: Optional[float] = None
    disparity_smoothness: float = 1e-3
    velocity_loss_scaling: Optional[float] = 0.05
    mask_dynamic: bool = False
    log_path: Path = Path('./log')
    save_frequency: int = -1
    save_val_depth: bool = False
    save_val_depth_batches: int = 0
    multiple_gpus: bool = False
    gpu_ids: Optional[Tuple[int, ...]] = None
    load_weights_folder: Optional[Path] = None
    use_wandb: Optional[bool] = False
