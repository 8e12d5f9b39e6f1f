"""Training-time augmentation (Loader.random_aug / Loader.augment / NYU.__getitem__, dataloader/loader.py:53-179, nyu_loader.py:38-66):
the numpy oracle (cv2-free restatement of cv2.warpPerspective / cv2.warpAffine) against items recorded from the unmodified reference
running the real cv2 (tests/golden/make_augment_golden.py), the host-side geometry / labels / random stream, and the device kernel
against both (GPU; bit-exact)."""
import os

import numpy as np
import pytest
import torch

from oracle import awr_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment_cases.npz"))
CASES = [(i, int(N), int(seed), int(D)) for i, (N, seed, D) in enumerate(G["meta"])]
OPS = ["trans", "scale", "rot", None]
KEYS = ["img", "jt_xyz", "jt_uvd", "center", "M", "cube"]
CUBE = np.asarray([300, 300, 300])                                   # NYU.cube (nyu_loader.py:16,24)


def _augs(i, N):
    return [(OPS[G[f"op{i}"][n]], G[f"trans{i}"][n], float(G[f"scale{i}"][n]), float(G[f"rot{i}"][n])) for n in range(N)]


@pytest.mark.parametrize("i,N,seed,D", CASES, ids=lambda v: str(v))
def test_oracle_matches_reference_items(i, N, seed, D):
    frames, jt_xyz, center_xyz = O.augment_case_inputs(N, seed)
    assert sorted(set(G[f"op{i}"].tolist())) == [0, 1, 2, 3]             # every augmentation branch is in the fixture
    for n, aug in enumerate(_augs(i, N)):
        out = O.nyu_train_item_np(frames[n], jt_xyz[n], center_xyz[n], CUBE, D, *aug)
        for o, k in zip(out, KEYS):
            assert o.dtype == np.float32 and np.array_equal(o, G[f"{k}{i}"][n]), (n, aug[0], k)      # bit-exact, cv2 warps included


@pytest.mark.parametrize("i,N,seed,D", CASES, ids=lambda v: str(v))
def test_random_stream_and_host_geometry_match_reference(i, N, seed, D):
    from awr_b200 import preprocess as PP
    rs = np.random.RandomState(23455)                                    # loader.py:10
    para = G[f"aug_para{i}"]
    frames, jt_xyz, center_xyz = O.augment_case_inputs(N, seed)
    for n, ref in enumerate(_augs(i, N)):
        op, trans, scale, rot = PP.random_aug(rs, *para)
        assert op == ref[0] and np.array_equal(trans, ref[1]) and scale == ref[2] and rot == ref[3]
        row, jx, ju, c, M, cube = PP.train_frame_geometry(jt_xyz[n], center_xyz[n], CUBE, D, O.NYU_PARAS, O.NYU_FLIP, (op, trans, scale, rot))
        for o, k in zip((jx, ju, c, M, cube), KEYS[1:]):
            assert o.dtype == np.float32 and np.array_equal(o, G[f"{k}{i}"][n]), (n, op, k)
        assert row.shape == (32,) and row[12] == {"trans": 1, "scale": 1, "rot": 2, None: 0}[op]


@pytest.mark.parametrize("i,N,seed,D", CASES, ids=lambda v: str(v))
def test_batched_host_geometry_is_bit_identical(i, N, seed, D):
    """preprocess.train_batch_geometry (array form, what train_batch uses) against the per-frame function and the reference's labels; plus a
    random stream of 64 further draws per case."""
    from awr_b200 import preprocess as PP
    frames, jt_xyz, center_xyz = O.augment_case_inputs(N, seed)
    bat = PP.train_batch_geometry(jt_xyz, center_xyz, CUBE, D, O.NYU_PARAS, O.NYU_FLIP, _augs(i, N))
    for o, k in zip(bat[1:], KEYS[1:]):
        assert o.dtype == np.float32 and np.array_equal(o, G[f"{k}{i}"]), k
    rs = np.random.RandomState(seed)
    _, jt2, c2 = O.augment_case_inputs(64, seed + 50)
    augs = [PP.random_aug(rs, *G[f"aug_para{i}"]) for _ in range(64)]
    augs[3] = ("trans", np.zeros(3), 1.0, 0.0)                           # the early-return branches (loader.py:109-110,167-168)
    augs[4] = ("scale", np.zeros(3), 1.0, 0.0)
    per = [PP.train_frame_geometry(jt2[n], c2[n], CUBE, D, O.NYU_PARAS, O.NYU_FLIP, augs[n]) for n in range(64)]
    bat = PP.train_batch_geometry(jt2, c2, CUBE, D, O.NYU_PARAS, O.NYU_FLIP, augs)
    for k in range(6):
        a = np.stack([p[k] for p in per])
        assert a.dtype == bat[k].dtype and np.array_equal(a, bat[k]), k


def test_matrix_helpers_match_cv2():
    cv2 = pytest.importorskip("cv2")
    from awr_b200 import preprocess as PP
    rng = np.random.RandomState(5)
    for t in range(300):
        s = 1 + 0.2 * rng.randn()
        M = np.array([[s, 0, rng.randn() * 30], [0, s, rng.randn() * 30], [0, 0, 1]], np.float32).astype(np.float64) if t % 2 else rng.randn(3, 3)
        assert np.array_equal(PP.invert3x3(M).reshape(3, 3), cv2.invert(M)[1]) and np.array_equal(O.cv_invert3x3_np(M), cv2.invert(M)[1])
        ang = rng.uniform(-360, 360)
        R = cv2.getRotationMatrix2D((64, 64), ang, 1)
        assert np.allclose(O.cv_rotation_matrix_2d_np((64, 64), ang), R, rtol=0, atol=1e-12)
        img = (rng.rand(64, 64) * 500 + 300).astype(np.float32)
        img[rng.rand(64, 64) < 0.4] = 0
        want = cv2.warpAffine(img, R, (64, 64), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(O.cv_warp_affine_linear_np(img, R), want)
        want = cv2.warpPerspective(img, M if t % 2 else np.eye(3) + 0.01 * M, (64, 64), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
        assert np.array_equal(O.cv_warp_perspective_linear_np(img, M if t % 2 else np.eye(3) + 0.01 * M), want)


@pytest.mark.gpu
@pytest.mark.parametrize("i,N,seed,D", CASES, ids=lambda v: str(v))
def test_device_train_batch_bit_exact(i, N, seed, D):
    from awr_b200 import preprocess as PP
    frames, jt_xyz, center_xyz = O.augment_case_inputs(N, seed)
    out = PP.train_batch(torch.from_numpy(frames).cuda(), jt_xyz, center_xyz, CUBE, D, O.NYU_PARAS, O.NYU_FLIP, _augs(i, N))
    assert out[0].shape == (N, 1, D, D) and out[0].is_cuda
    for o, k in zip(out, KEYS):
        got = o.cpu().numpy()
        bad = [n for n in range(N) if not np.array_equal(got[n], G[f"{k}{i}"][n])]
        assert not bad, (k, [(n, OPS[G[f"op{i}"][n]]) for n in bad])
    # the NYU wire format (nyu_loader.py:71-74)
    d16 = frames.astype(np.uint16)
    bgr = np.stack([(d16 & 255).astype(np.uint8), (d16 >> 8).astype(np.uint8), np.zeros_like(d16, dtype=np.uint8)], axis=-1)
    out2 = PP.train_batch(torch.from_numpy(bgr).cuda(), jt_xyz, center_xyz, CUBE, D, O.NYU_PARAS, O.NYU_FLIP, _augs(i, N))
    assert torch.equal(out2[0], out[0])
    with pytest.raises(ValueError):
        PP.train_batch(torch.from_numpy(frames).cuda(), jt_xyz, center_xyz, CUBE, D, O.NYU_PARAS, O.NYU_FLIP, _augs(i, N)[:-1])
