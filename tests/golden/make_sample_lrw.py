"""Extract the reference's own real-input fixture (Datasets/SAMPLE_LRW: 10 clips of the word ABOUT) into a compact form
that can travel to the GPU box: mouth crops as uint8 [10,29,96,96,3] (RGB, decoded exactly as datasets/lrw/dataset.py:20-24
does) and the 16 kHz audio [10,19456].  Build container only (needs /root/reference and cv2).

    python tests/golden/make_sample_lrw.py
"""
import bz2
import glob
import os
import pickle

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = "/root/reference/Datasets/SAMPLE_LRW"


def loadframes(filename):                       # same decode as the reference loader
    with bz2.BZ2File(filename, "r") as f:
        data = pickle.load(f)
    return np.array([cv2.imdecode(im, cv2.IMREAD_COLOR)[:, :, ::-1] for im in data])


def main():
    mouths = sorted(glob.glob(os.path.join(ROOT, "LRW_Faces", "*", "test", "*_mouth.npz")))
    assert len(mouths) == 10, mouths
    video, audio, names = [], [], []
    for m in mouths:
        rel = os.path.relpath(m, os.path.join(ROOT, "LRW_Faces"))[: -len("_mouth.npz")]
        a = np.load(os.path.join(ROOT, "lipread_audio", rel + ".npz"))["data"].astype(np.float32)
        fr = loadframes(m)
        assert fr.shape == (29, 96, 96, 3) and fr.dtype == np.uint8, fr.shape
        video.append(fr); audio.append(a); names.append(rel)
    video = np.stack(video); audio = np.stack(audio)
    out = os.path.join(HERE, "sample_lrw.npz")
    np.savez_compressed(out, video=video, audio=audio, names=np.array(names))
    print(video.shape, audio.shape, os.path.getsize(out))


if __name__ == "__main__":
    main()
