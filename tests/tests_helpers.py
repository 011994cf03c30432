"""Shared fixtures for tests that read the SAMPLE_LRW clips (tests/golden/sample_lrw.npz)."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def load_sample_lrw():
    d = np.load(os.path.join(HERE, "golden", "sample_lrw.npz"))
    v = d["video"].astype(np.float32) / 255.0                       # datasets/lrw/dataset.py:82-86
    v = (v - MEAN) / STD
    video = torch.from_numpy(v).permute(0, 4, 1, 2, 3).contiguous()  # [B,3,T,H,W]
    return video, torch.from_numpy(d["audio"])
