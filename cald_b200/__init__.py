"""cald_b200 -- B200-native engine for CALD's unlabeled-pool consistency scoring.

Public surface (mirrors we1pingyu/CALD cald_train.py):
    get_uncertainty(task_model, unlabeled_loader, augs, num_cls) -> (consistency_all, cls_all)
    select(uncertainty, cls_corrs, subset, labeled_loader, budget_num, ...)
    cls_kldiv(labeled_loader, cls_corrs, budget, cycle)
and, on the same engine, the detection-only baselines and the evaluation forward:
    lt_c_uncertainty(task_model, unlabeled_loader)          (lt_c_train.py:105-121)
    ls_c_uncertainty(task_model, unlabeled_loader)          (ls_c_train.py:108-155)
    EngineModel(task_model)(images) -> [detections dict]     (detection/engine.py evaluation loops)
Everything below these calls runs in hand-written sm_100a CUDA behind libcald_b200.so.
"""
from .api import (get_uncertainty, select, cls_kldiv, score_images, engine_for, lt_c_uncertainty,  # noqa: F401
                  close_engines, get_uncertainty_files,
                  ls_c_uncertainty, EngineModel)
from .engine import Engine  # noqa: F401

__all__ = ["get_uncertainty", "select", "cls_kldiv", "score_images", "engine_for", "Engine", "lt_c_uncertainty",
           "ls_c_uncertainty", "EngineModel", "close_engines", "get_uncertainty_files"]
