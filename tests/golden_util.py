"""Access to the committed golden fixtures (outputs of the unmodified reference)."""
import os

import numpy as np
import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attack_golden.npz")
GOLD = dict(np.load(_PATH))


def T(key: str) -> torch.Tensor:
    return torch.from_numpy(np.array(GOLD[key]))


def text(key: str) -> str:
    return bytes(GOLD[key].tobytes()).decode()
