"""Device-side depth preprocessing of the reference's non-augmented data path (test / validation):
`Loader.crop` + `Loader.normalize` (dataloader/loader.py:19-51,88-101) for a batch of raw frames in one launch of csrc/preprocess.cu.

The per-frame box geometry (`center2bounds`, `center2transmat`: a dozen float64 scalars) is computed here on the host exactly as the
reference does; every pixel operation (box gather with zero padding, cube clamp, cv2.resize INTER_NEAREST index rule, centring pad,
max-depth / invalid -> background, clip, scale to [-1,1]) runs on the GPU.  The result is bit-identical to the reference's numpy/cv2 code.
"""
import numpy as np
import torch

from . import _lib as L


def center2bounds(center, csize, paras):
    """loader.py:181-188."""
    center, csize, p2 = np.asarray(center), np.asarray(csize, dtype=np.float64), np.asarray(paras[:2])
    ustart, vstart = center[:2] - (csize[:2] / 2.) / center[2] * p2 + 0.5
    uend, vend = center[:2] + (csize[:2] / 2.) / center[2] * p2 + 0.5
    return int(ustart), int(uend), int(vstart), int(vend), center[2] - csize[2] / 2., center[2] + csize[2] / 2.


def center2transmat(center, csize, dsize, paras):
    """loader.py:210-240."""
    ustart, uend, vstart, vend, _, _ = center2bounds(center, csize, paras)
    trans1 = np.eye(3); trans1[0][2] = -ustart; trans1[1][2] = -vstart
    w, h = (uend - ustart), (vend - vstart)
    scale = min(dsize[0] / w, dsize[1] / h)
    size = (int(w * scale), int(h * scale))
    sc = scale * np.eye(3); sc[2][2] = 1
    trans2 = np.eye(3)
    trans2[0][2] = int(np.floor(dsize[0] / 2. - size[0] / 2.)); trans2[1][2] = int(np.floor(dsize[1] / 2. - size[1] / 2.))
    return np.dot(trans2, np.dot(sc, trans1)).astype(np.float32)


def crop_params(center_uvd, center_z, cube, img_size, paras):
    """(N,12) float64 parameter block of awr_crop_normalize + the (N,3,3) float32 crop affines."""
    center_uvd, cube = np.asarray(center_uvd, dtype=np.float32), np.asarray(cube, dtype=np.float64)
    N = center_uvd.shape[0]
    dsize = np.array([img_size, img_size])
    P, Ms = np.zeros((N, 12), np.float64), np.zeros((N, 3, 3), np.float32)
    for n in range(N):
        ustart, uend, vstart, vend, zstart, zend = center2bounds(center_uvd[n], cube[n], paras)
        w, h = (uend - ustart), (vend - vstart)
        if w <= 0 or h <= 0:
            raise ValueError(f"frame {n}: empty crop box (centre depth {center_uvd[n][2]})")
        scale = min(dsize[0] / w, dsize[1] / h)
        size = (int(w * scale), int(h * scale))
        us, vs = (dsize - size) / 2.
        P[n] = [ustart, vstart, w, h, size[0], size[1], int(us), int(vs), zstart, zend, np.float64(center_z[n]), cube[n][2] / 2.]
        Ms[n] = center2transmat(center_uvd[n], cube[n], dsize, paras)
    return P, Ms


def crop_normalize(frames, center_uvd, center_z, cube, img_size, paras):
    """frames: CUDA tensor (N,Hs,Ws) float32 millimetres, or (N,Hs,Ws,3) uint8 BGR as cv2.imread returns the NYU PNGs (nyu_loader.py:71-74).
    center_uvd (N,3): hand centre in image coordinates (u, v, depth mm); center_z (N,): z of center_xyz (nyu_loader.py:60); cube (N,3) mm.
    Returns (img (N,1,img_size,img_size) float32 CUDA, M (N,3,3) float32 CPU) = what Loader.crop + Loader.normalize give per frame."""
    L.require_cuda(frames)
    if frames.dtype == torch.float32 and frames.dim() == 3:
        fmt = 0
    elif frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3:
        fmt = 1
    else:
        raise ValueError("frames must be (N,Hs,Ws) float32 or (N,Hs,Ws,3) uint8")
    frames = frames.contiguous()
    N, Hs, Ws = frames.shape[:3]
    P, Ms = crop_params(center_uvd, center_z, cube, img_size, paras)
    params = torch.from_numpy(P).to(frames.device)
    out = torch.empty(N, 1, img_size, img_size, dtype=torch.float32, device=frames.device)
    L.check(L.lib().awr_crop_normalize(L.ptr(frames), fmt, N, Hs, Ws, L.ptr(params), int(img_size), L.ptr(out), L.stream()), "awr_crop_normalize")
    return out, torch.from_numpy(Ms)
