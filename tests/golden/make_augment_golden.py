"""Records what the UNMODIFIED reference training data path -- dataloader/nyu_loader.py::NYU.__getitem__ (phase 'train': Loader.crop ->
Loader.random_aug -> Loader.augment (translate / rotate / scale via the real cv2.warpPerspective / cv2.warpAffine) -> labels) -- produces
on seeded synthetic raw frames -> tests/golden/augment_cases.npz.  The NYU object is built without its dataset files (object.__new__ +
the attributes NYU.__init__ sets) and `nyu_reader` is pointed at the synthetic frames; every method that runs is the reference's own.
Run in the build container (needs /root/reference, cv2, scipy)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
from dataloader.nyu_loader import NYU         # noqa: E402  (the reference class)
from dataloader.loader import Loader          # noqa: E402
from oracle import awr_oracle as O            # noqa: E402  (input generator only)


def case(N, seed, img_size, aug_para):
    frames, jt_xyz, center_xyz = O.augment_case_inputs(N, seed)
    ds = object.__new__(NYU)
    Loader.__init__(ds, "", "train", img_size, "nyu")              # RandomState(23455), aug_ops
    ds.name, ds.val, ds.paras, ds.flip = "nyu", False, (588.03, 587.07, 320., 240.), -1
    ds.cube, ds.dsize, ds.img_size, ds.jt_num, ds.aug_para = np.asarray([300, 300, 300]), np.asarray([img_size, img_size]), img_size, 14, aug_para
    ds.data = [(n, None, jt_xyz[n], center_xyz[n]) for n in range(N)]
    ds.nyu_reader = lambda idx: frames[idx].copy()
    draws = []
    ref_random_aug = ds.random_aug

    def recording_random_aug(*a):
        d = ref_random_aug(*a)
        draws.append(d)
        return d
    ds.random_aug = recording_random_aug
    items = [ds[n] for n in range(N)]
    ops = np.array([{"trans": 0, "scale": 1, "rot": 2, None: 3}[d[0]] for d in draws])
    return dict(N=N, seed=seed, img_size=img_size, aug_para=np.array(aug_para, dtype=np.float64), op=ops,
                trans=np.stack([d[1] for d in draws]), scale=np.array([d[2] for d in draws]), rot=np.array([d[3] for d in draws]),
                img=np.stack([it[0] for it in items]), jt_xyz=np.stack([it[1] for it in items]), jt_uvd=np.stack([it[2] for it in items]),
                center=np.stack([it[3] for it in items]), M=np.stack([it[4] for it in items]), cube=np.stack([it[5] for it in items]))


if __name__ == "__main__":
    cases = [case(16, 21, 128, [10, 0.1, 180]), case(12, 22, 256, [35., 0.05, 180.]), case(12, 23, 96, [10, 0.1, 180])]
    flat = {"meta": np.array([[c["N"], c["seed"], c["img_size"]] for c in cases])}
    for i, c in enumerate(cases):
        for k in ("aug_para", "op", "trans", "scale", "rot", "img", "jt_xyz", "jt_uvd", "center", "M", "cube"):
            flat[f"{k}{i}"] = c[k]
    np.savez_compressed(os.path.join(HERE, "augment_cases.npz"), **flat)
    for c in cases:
        print(c["N"], c["img_size"], "ops", np.bincount(c["op"], minlength=4), "fg fraction %.3f" % float((c["img"] < 0.99).mean()))
