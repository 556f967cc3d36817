"""Checkpoint / seeding helpers with the reference's names and file formats (row a14).

Format A (``mnist/utils.py:16-31``): ``{'epoch', 'model_state_dict', 'optimizer_state_dict'}``;
Format B (``src/utils/tools.py:17-29``): a raw ``state_dict``.  ``load_checkpoint`` accepts both.
"""
import os
import random
from pathlib import Path

import numpy as np
import torch


def set_seed(seed):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def save_checkpoint(model, optimizer, epoch, path):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict() if optimizer else {}}, path)


def load_checkpoint(model, optimizer, path, device):
    ckpt = torch.load(path, map_location=device)
    if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
        model.load_state_dict(ckpt["model_state_dict"])
        if optimizer:
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        return ckpt.get("epoch", 0)
    model.load_state_dict(ckpt)
    return 0


class CheckpointManager:
    def __init__(self, base_dir, exp_name, run_id):
        self.base_dir = Path(base_dir) / exp_name / run_id

    def get_path(self, type="checkpoints"):
        path = self.base_dir / type
        path.mkdir(parents=True, exist_ok=True)
        return path

    def _file(self, model_name, epoch):
        name = f"{model_name}_final.pth" if epoch is None else f"{model_name}_epoch_{epoch}.pth"
        return self.get_path("checkpoints") / name

    def save(self, model, model_name, epoch=None):
        torch.save(model.state_dict(), self._file(model_name, epoch))

    def load(self, model, model_name, device, epoch=None):
        path = self._file(model_name, epoch)
        if not path.exists():
            raise FileNotFoundError(f"Checkpoint {path} not found.")
        model.load_state_dict(torch.load(path, map_location=device))
        return model
