"""Depth preprocessing (Loader.crop + Loader.normalize, dataloader/loader.py): the numpy oracle against frames recorded from the
unmodified reference running the real cv2 (CPU), the host-side box geometry, and the device kernel against both (GPU; bit-exact)."""
import os

import numpy as np
import pytest
import torch

from oracle import awr_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_cases.npz"))
CASES = [(i, int(N), int(seed), int(D)) for i, (N, seed, D) in enumerate(G["meta"])]


@pytest.mark.parametrize("i,N,seed,D", CASES, ids=lambda v: str(v))
def test_oracle_matches_reference_frames(i, N, seed, D):
    frames, centers, cubes = O.preprocess_case_inputs(N, seed)
    outs = [O.crop_normalize_np(frames[n], centers[n], np.float64(centers[n][2]), cubes[n], D) for n in range(N)]
    assert np.array_equal(np.stack([o[0] for o in outs]), G[f"img{i}"])          # bit-exact, cv2.resize index rule included
    assert np.array_equal(np.stack([o[1] for o in outs]), G[f"M{i}"])


@pytest.mark.parametrize("i,N,seed,D", CASES, ids=lambda v: str(v))
def test_host_geometry_matches_reference(i, N, seed, D):
    from awr_b200 import preprocess as PP
    frames, centers, cubes = O.preprocess_case_inputs(N, seed)
    P, Ms = PP.crop_params(centers, centers[:, 2].astype(np.float64), cubes, D, O.NYU_PARAS)
    assert np.array_equal(Ms, G[f"M{i}"])
    for n in range(N):
        b = O.center2bounds_np(centers[n], cubes[n])
        assert (P[n, 0], P[n, 1], P[n, 2], P[n, 3]) == (b[0], b[2], b[1] - b[0], b[3] - b[2]) and P[n, 8] == b[4] and P[n, 9] == b[5]


@pytest.mark.gpu
@pytest.mark.parametrize("i,N,seed,D", CASES, ids=lambda v: str(v))
def test_device_crop_normalize_bit_exact(i, N, seed, D):
    from awr_b200 import preprocess as PP
    frames, centers, cubes = O.preprocess_case_inputs(N, seed)
    cz = centers[:, 2].astype(np.float64)
    img, M = PP.crop_normalize(torch.from_numpy(frames).cuda(), centers, cz, cubes, D, O.NYU_PARAS)
    assert img.shape == (N, 1, D, D) and np.array_equal(img.cpu().numpy()[:, 0], G[f"img{i}"]) and np.array_equal(M.numpy(), G[f"M{i}"])
    # the NYU wire format (nyu_loader.py:71-74): 16-bit depth split over the B and G bytes of a BGR frame
    d16 = frames.astype(np.uint16)
    bgr = np.stack([(d16 & 255).astype(np.uint8), (d16 >> 8).astype(np.uint8), np.zeros_like(d16, dtype=np.uint8)], axis=-1)
    img2, _ = PP.crop_normalize(torch.from_numpy(bgr).cuda(), centers, cz, cubes, D, O.NYU_PARAS)
    assert torch.equal(img2, img)
    with pytest.raises(ValueError):
        PP.crop_normalize(torch.zeros(2, 8, 8, dtype=torch.float64, device="cuda"), centers[:2], cz[:2], cubes[:2], D, O.NYU_PARAS)
