"""Records what the UNMODIFIED reference dataloader/loader.py::Loader.crop + Loader.normalize (with the real cv2) produce on seeded
synthetic raw frames -> tests/golden/preprocess_cases.npz (compressed: most pixels are background).  Run in the build container (needs /root/reference and cv2)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
from dataloader.loader import Loader         # noqa: E402  (the reference class)
from oracle import awr_oracle as O           # noqa: E402  (input generator only)


def case(N, seed, img_size):
    frames, centers, cubes = O.preprocess_case_inputs(N, seed)
    ld = Loader("", "test", img_size, "nyu")
    ld.paras, ld.flip = O.NYU_PARAS, O.NYU_FLIP
    imgs, Ms = [], []
    for n in range(N):
        img, M = ld.crop(frames[n].copy(), centers[n], cubes[n], np.array([img_size, img_size]))
        center_xyz_z = np.float64(centers[n][2])                 # nyu_loader.py:60 passes center_xyz; its z equals the uvd depth
        img = ld.normalize(img.max(), img, np.array([0.0, 0.0, center_xyz_z]), cubes[n])
        imgs.append(img.astype(np.float32)); Ms.append(M.astype(np.float32))
    return dict(N=N, seed=seed, img_size=img_size, img=np.stack(imgs), M=np.stack(Ms))


if __name__ == "__main__":
    cases = [case(8, 3, 128), case(5, 4, 256), case(4, 5, 96)]
    flat = {"meta": np.array([[c["N"], c["seed"], c["img_size"]] for c in cases])}
    for i, c in enumerate(cases):
        flat[f"img{i}"], flat[f"M{i}"] = c["img"], c["M"]
    np.savez_compressed(os.path.join(HERE, "preprocess_cases.npz"), **flat)
    for c in cases:
        print(c["N"], c["img_size"], "fg fraction %.3f" % float((c["img"] < 0.99).mean()), "min %.3f" % c["img"].min())
