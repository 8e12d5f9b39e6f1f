"""Device-side depth preprocessing of the reference's data path, a batch of raw frames per launch of
csrc/preprocess.cu: `Loader.crop` + `Loader.normalize` (dataloader/loader.py:19-51,88-101) for test / validation, and the training path
`NYU.__getitem__` (nyu_loader.py:38-66) with `Loader.random_aug` / `Loader.augment` (loader.py:53-86: translate / scale through
cv2.warpPerspective, rotate through cv2.warpAffine, :103-179) -- `train_batch`.

The per-frame box geometry (`center2bounds`, `center2transmat`: a dozen float64 scalars) is computed here on the host exactly as the
reference does; every pixel operation (box gather with zero padding, cube clamp, cv2.resize INTER_NEAREST index rule, centring pad,
max-depth / invalid -> background, clip, scale to [-1,1]) runs on the GPU.  The result is bit-identical to the reference's numpy/cv2 code.
"""
import math

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


# ---- training path: random augmentation (loader.py:53-179, nyu_loader.py:38-66) ---------------------------------------------------------
AUG_OPS = ("trans", "scale", "rot", None)          # loader.py:17: one of these is drawn per frame


def uvd2xyz(pts, paras, flip):
    """util/util.py:13-20."""
    q = np.array(pts, copy=True).reshape(-1, 3)
    q[:, :2] = (q[:, :2] - np.asarray(paras[2:])) * q[:, 2:] / np.asarray(paras[:2])
    q[:, 1] *= flip
    return q.reshape(np.shape(pts)).astype(np.float32)


def xyz2uvd(pts, paras, flip):
    """util/util.py:3-10."""
    q = np.array(pts, copy=True).reshape(-1, 3)
    q[:, 1] *= flip
    q[:, :2] = q[:, :2] * np.asarray(paras[:2]) / q[:, 2:] + np.asarray(paras[2:])
    return q.reshape(np.shape(pts)).astype(np.float32)


def random_aug(rs, sigma_trans=None, sigma_scale=None, sigma_rot=None):
    """Loader.random_aug (loader.py:53-72) on a numpy RandomState: the same draws in the same order, so a loader seeded like the reference's
    (RandomState(23455), loader.py:10) produces the reference's augmentation stream.  Returns (op, trans (3,), scale, rot degrees)."""
    sigma_trans = 35. if sigma_trans is None else sigma_trans
    sigma_scale = 0.05 if sigma_scale is None else sigma_scale
    sigma_rot = 180. if sigma_rot is None else sigma_rot
    op = AUG_OPS[rs.randint(0, len(AUG_OPS))]
    trans = rs.randn(3) * sigma_trans
    scale = abs(1. + rs.randn() * sigma_scale)
    rot = rs.uniform(-sigma_rot, sigma_rot)
    return op, trans, scale, rot


def invert3x3(M):
    """What cv2.warpPerspective does to its matrix first (cv2.invert, closed form for 3x3: adjugate / determinant, float64)."""
    a, b, c, d, e, f, g, h, i = (float(v) for v in np.asarray(M, dtype=np.float64).ravel())
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
    if det == 0.0:
        return np.zeros(9)
    r = 1.0 / det
    return np.array([(e * i - f * h) * r, (c * h - b * i) * r, (b * f - c * e) * r, (f * g - d * i) * r, (a * i - c * g) * r, (c * d - a * f) * r,
                     (d * h - e * g) * r, (b * g - a * h) * r, (a * e - b * d) * r])


def rotation_inverse_map(size_hw, angle_deg):
    """cv2.getRotationMatrix2D((w//2, h//2), angle, 1) followed by the inversion cv2.warpAffine applies to a forward 2x3 map (float64)."""
    cx, cy = size_hw[1] // 2, size_hw[0] // 2
    a = angle_deg * math.pi / 180.0
    al, be = math.cos(a), math.sin(a)
    m = [al, be, (1 - al) * cx - be * cy, -be, al, be * cx + (1 - al) * cy]
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = m[4] * D, m[0] * D
    m[0], m[1], m[3], m[4] = a11, m[1] * -D, m[3] * -D, a22
    b1, b2 = -m[0] * m[2] - m[1] * m[5], -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return np.array(m)


def _perspective_tile_width(D):
    bh = min(16, D)
    bw = min(1024 // bh, D)
    return bw


def rotate_pts(pt, center, angle):
    """loader.py:242-252."""
    alpha = angle * np.pi / 180.
    r = pt.copy()
    r[:, 0] = (pt[:, 0] - center[0]) * np.cos(alpha) - (pt[:, 1] - center[1]) * np.sin(alpha)
    r[:, 1] = (pt[:, 0] - center[0]) * np.sin(alpha) + (pt[:, 1] - center[1]) * np.cos(alpha)
    r[:, :2] += center[:2]
    return r.astype(np.float32)


def train_frame_geometry(jt_xyz, center_xyz, cube, img_size, paras, flip, aug):
    """Host half of NYU.__getitem__ (train phase) for one frame: everything that is O(joints) -- the crop box, the augmentation's effect
    on centre / cube / crop affine / joint labels -- plus the 32-double parameter row of awr_crop_augment_normalize for the pixels.
    aug = (op, trans, scale, rot) as random_aug returns it.  Returns (row (32,) f64, jt_xyz_norm, jt_uvd_norm, center_xyz, M, cube) with the
    dtypes NYU.__getitem__ returns (float32)."""
    op, trans, scale, rot = aug
    cube = np.asarray(cube)
    shape = np.array([img_size, img_size])
    center = xyz2uvd(np.asarray(center_xyz, dtype=np.float64), paras, flip)
    jt = np.asarray(jt_xyz, dtype=np.float64) - center_xyz
    row = np.zeros(32, np.float64)
    ustart, uend, vstart, vend, zstart, zend = center2bounds(center, cube, paras)
    w, h = (uend - ustart), (vend - vstart)
    if w <= 0 or h <= 0:
        raise ValueError(f"empty crop box (centre depth {center[2]})")
    sc = min(shape[0] / w, shape[1] / h)
    size = (int(w * sc), int(h * sc))
    us, vs = (shape - size) / 2.
    row[:10] = [ustart, vstart, w, h, size[0], size[1], int(us), int(vs), zstart, zend]
    M = center2transmat(center, cube, shape, paras)
    if op == "trans" and not np.allclose(trans, 0.):                       # Loader.translate, loader.py:103-123
        new_center = xyz2uvd(uvd2xyz(center, paras, flip) + trans, paras, flip)
        if not np.allclose(center[2], 0.) or np.allclose(new_center[2], 0.):
            new_M = center2transmat(new_center, cube, shape, paras)
            row[12] = 1
            row[13:22] = invert3x3(np.dot(new_M, np.linalg.inv(M)))
            row[22:24] = center2bounds(new_center, cube, paras)[4:]
        else:
            new_M = M
        jt = jt + uvd2xyz(center, paras, flip) - uvd2xyz(new_center, paras, flip)
        center, M = new_center, new_M
    elif op == "rot":                                                      # Loader.rotate, loader.py:141-161
        r = np.mod(rot, 360)
        row[12] = 2
        row[13:19] = rotation_inverse_map((img_size, img_size), -r)
        c_xyz = uvd2xyz(center, paras, flip)
        jt = uvd2xyz(rotate_pts(xyz2uvd(jt + c_xyz, paras, flip), center, r), paras, flip) - c_xyz
    elif op == "scale" and not np.allclose(scale, 1.):                     # Loader.scale, loader.py:163-179
        new_cube = cube * scale
        if not np.allclose(center[2], 0.):
            new_M = center2transmat(center, new_cube, shape, paras)
            row[12] = 1
            row[13:22] = invert3x3(np.dot(new_M, np.linalg.inv(M)))
            row[22:24] = center2bounds(center, new_cube, paras)[4:]
        else:
            new_M = M
        cube, M = new_cube, new_M
    row[24] = _perspective_tile_width(img_size)
    row[10], row[11] = center[2], cube[2] / 2.                             # Loader.normalize's centre z and half cube (after augmentation)
    c_xyz = uvd2xyz(center, paras, flip)                                   # nyu_loader.py:58-66: the labels
    q = xyz2uvd(jt + c_xyz, paras, flip)
    hom = np.dot(M, np.hstack([q[:, :2], np.ones((q.shape[0], 1))]).T).T
    hom[:, :2] /= hom[:, 2:]
    jt_uvd = np.hstack([hom[:, :2], q[:, 2:]]).astype(np.float32)
    jt_uvd[:, :2] = jt_uvd[:, :2] / (img_size / 2.) - 1
    jt_uvd[:, 2] = (jt_uvd[:, 2] - c_xyz[2]) / (cube[2] / 2.0)
    return (row, (jt / (cube / 2.)).astype(np.float32), jt_uvd.astype(np.float32), c_xyz.astype(np.float32), M.astype(np.float32),
            cube.astype(np.float32))


def _bounds_batch(center, csize, paras):
    """center2bounds for N frames: center (N,3) float32, csize (N,3) float64.  Same float64 expressions, element by element."""
    p2 = np.asarray(paras[:2])
    half = (csize[:, :2] / 2.) / center[:, 2:3] * p2
    lo, hi = center[:, :2] - half + 0.5, center[:, :2] + half + 0.5
    z = center[:, 2].astype(np.float64)
    return (np.trunc(lo[:, 0]).astype(np.int64), np.trunc(hi[:, 0]).astype(np.int64), np.trunc(lo[:, 1]).astype(np.int64),
            np.trunc(hi[:, 1]).astype(np.int64), z - csize[:, 2] / 2., z + csize[:, 2] / 2.)


def _transmat_batch(ustart, uend, vstart, vend, D):
    """center2transmat for N frames from their bounds, in closed form: trans2 . diag(sc, sc, 1) . trans1 has one rounded product and one
    addition per translation entry, which is what the 3x3 float64 matrix products of the reference evaluate to.  Returns
    (M (N,3,3) float32, sc, size_w, size_h, t2x, t2y)."""
    w, h = uend - ustart, vend - vstart
    sc = np.minimum(D / w, D / h)
    sw, sh = np.trunc(w * sc).astype(np.int64), np.trunc(h * sc).astype(np.int64)
    t2x, t2y = np.floor(D / 2. - sw / 2.).astype(np.int64), np.floor(D / 2. - sh / 2.).astype(np.int64)
    M = np.zeros((len(w), 3, 3), np.float64)
    M[:, 0, 0] = sc; M[:, 1, 1] = sc; M[:, 2, 2] = 1.0
    M[:, 0, 2] = sc * (-ustart) + t2x
    M[:, 1, 2] = sc * (-vstart) + t2y
    return M.astype(np.float32), sc, sw, sh, t2x, t2y


def train_batch_geometry(jt_xyz, center_xyz, cube, img_size, paras, flip, augs):
    """train_frame_geometry for a whole batch: the same float32 / float64 expressions evaluated on arrays with a leading frame axis (grouped
    by augmentation, because the rotate branch continues in float32 where the others stay in float64), with only the calls whose rounding
    depends on the library routine left per frame (the float32 LAPACK inverse and 3x3 products, libm cos / sin).  Bit-identical to the
    per-frame function (tests/test_augment.py); ~10x faster.  Returns (rows (N,32) f64, jt_xyz_norm (N,J,3), jt_uvd_norm (N,J,3),
    center_xyz (N,3), M (N,3,3), cube (N,3)), float32 labels."""
    N = len(augs)
    D = int(img_size)
    jt_xyz, center_xyz = np.asarray(jt_xyz, dtype=np.float64), np.asarray(center_xyz, dtype=np.float64)
    J = jt_xyz.shape[1]
    cube0 = np.asarray(cube)
    csize = np.broadcast_to(np.asarray(cube0, dtype=np.float64), (N, 3)).copy()              # per frame; scaled below
    center = xyz2uvd(center_xyz, paras, flip)                                                 # (N,3) float32
    jt64 = jt_xyz - center_xyz[:, None, :]
    rows = np.zeros((N, 32), np.float64)
    us, ue, vs, ve, zs, ze = _bounds_batch(center, csize, paras)
    if np.any(ue - us <= 0) or np.any(ve - vs <= 0):
        n = int(np.argmax((ue - us <= 0) | (ve - vs <= 0)))
        raise ValueError(f"frame {n}: empty crop box (centre depth {center[n][2]})")
    M, sc, sw, sh, t2x, t2y = _transmat_batch(us, ue, vs, ve, D)
    rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 4], rows[:, 5] = us, vs, ue - us, ve - vs, sw, sh
    rows[:, 6], rows[:, 7] = np.trunc((D - sw) / 2.), np.trunc((D - sh) / 2.)
    rows[:, 8], rows[:, 9] = zs, ze
    ops = [a[0] for a in augs]
    trans = np.stack([np.asarray(a[1], dtype=np.float64) for a in augs])
    scale = np.array([float(a[2]) for a in augs])
    is_trans = np.array([o == "trans" for o in ops]) & ~np.all(np.isclose(trans, 0.), axis=1)
    is_rot = np.array([o == "rot" for o in ops])
    is_scale = np.array([o == "scale" for o in ops]) & ~np.isclose(scale, 1.)
    new_M = M.copy()
    warp = np.zeros(N, bool)                                                                  # frames that go through Loader.recrop

    # ---- translate (loader.py:103-123) ------------------------------------------------------------------------------------------------
    if is_trans.any():
        i = np.nonzero(is_trans)[0]
        c_old = center[i]
        c_new = xyz2uvd(uvd2xyz(c_old, paras, flip) + trans[i], paras, flip)
        redo = ~np.isclose(c_old[:, 2], 0.) | np.isclose(c_new[:, 2], 0.)
        b = _bounds_batch(c_new, csize[i], paras)
        Mn = _transmat_batch(b[0], b[1], b[2], b[3], D)[0]
        new_M[i[redo]] = Mn[redo]
        warp[i[redo]] = True
        rows[i[redo], 22], rows[i[redo], 23] = b[4][redo], b[5][redo]
        jt64[i] = jt64[i] + uvd2xyz(c_old, paras, flip)[:, None, :] - uvd2xyz(c_new, paras, flip)[:, None, :]
        center[i] = c_new
    # ---- scale (loader.py:163-179) ------------------------------------------------------------------------------------------------------
    if is_scale.any():
        i = np.nonzero(is_scale)[0]
        csize[i] = cube0 * scale[i][:, None]
        redo = ~np.isclose(center[i][:, 2], 0.)
        b = _bounds_batch(center[i], csize[i], paras)
        Mn = _transmat_batch(b[0], b[1], b[2], b[3], D)[0]
        new_M[i[redo]] = Mn[redo]
        warp[i[redo]] = True
        rows[i[redo], 22], rows[i[redo], 23] = b[4][redo], b[5][redo]
    if warp.any():
        inv = np.linalg.inv(M[warp])                                                          # float32 LAPACK inverse per matrix, as the reference calls it
        for k, n in enumerate(np.nonzero(warp)[0]):
            rows[n, 12] = 1
            rows[n, 13:22] = invert3x3(np.dot(new_M[n], inv[k]))
    M = new_M
    # ---- rotate (loader.py:141-161): continues in float32 -------------------------------------------------------------------------------
    jt32 = None
    if is_rot.any():
        i = np.nonzero(is_rot)[0]
        r = [np.mod(augs[n][3], 360) for n in i]
        rows[i, 12] = 2
        for k, n in enumerate(i):
            rows[n, 13:19] = rotation_inverse_map((D, D), -r[k])
        alpha = [rk * np.pi / 180. for rk in r]
        ca, sa = np.array([np.cos(a) for a in alpha])[:, None], np.array([np.sin(a) for a in alpha])[:, None]
        c = center[i]
        c_xyz = uvd2xyz(c, paras, flip)
        pt = xyz2uvd(jt64[i] + c_xyz[:, None, :], paras, flip)                                # (n,J,3) float32
        rot = pt.copy()
        dx, dy = pt[:, :, 0] - c[:, 0:1], pt[:, :, 1] - c[:, 1:2]
        rot[:, :, 0] = dx * ca - dy * sa
        rot[:, :, 1] = dx * sa + dy * ca
        rot[:, :, :2] += c[:, None, :2]
        jt32 = uvd2xyz(rot.astype(np.float32), paras, flip) - c_xyz[:, None, :]             # float32 from here on
    rows[:, 24] = _perspective_tile_width(D)
    rows[:, 10], rows[:, 11] = center[:, 2], csize[:, 2] / 2.
    # ---- labels (nyu_loader.py:58-66) ---------------------------------------------------------------------------------------------------
    c_xyz = uvd2xyz(center, paras, flip)
    q = np.empty((N, J, 3), np.float32)
    jt_n = np.empty((N, J, 3), np.float32)
    f64 = ~is_rot
    q[f64] = xyz2uvd(jt64[f64] + c_xyz[f64][:, None, :], paras, flip)
    jt_n[f64] = (jt64[f64] / (csize[f64] / 2.)[:, None, :]).astype(np.float32)
    if jt32 is not None:
        q[is_rot] = xyz2uvd(jt32 + c_xyz[is_rot][:, None, :], paras, flip)
        jt_n[is_rot] = (jt32 / (csize[is_rot] / 2.)[:, None, :]).astype(np.float32)
    pts = np.concatenate([q[:, :, :2], np.ones((N, J, 1))], axis=2)                           # (N,J,3) float64, as the reference's hstack
    hom = np.stack([np.dot(M[n], pts[n].T).T for n in range(N)])                              # float32 x float64 3x3 products: per frame, as the reference
    hom[:, :, :2] /= hom[:, :, 2:]
    jt_uvd = np.concatenate([hom[:, :, :2], q[:, :, 2:]], axis=2).astype(np.float32)
    jt_uvd[:, :, :2] = jt_uvd[:, :, :2] / (D / 2.) - 1
    jt_uvd[:, :, 2] = (jt_uvd[:, :, 2] - c_xyz[:, 2:3]) / (csize[:, 2:3] / 2.0)
    return rows, jt_n, jt_uvd, c_xyz.astype(np.float32), M.astype(np.float32), csize.astype(np.float32)


def train_batch(frames, jt_xyz, center_xyz, cube, img_size, paras, flip, augs):
    """The reference's training items for a batch of raw frames (what collating NYU.__getitem__ over N indices returns), pixels on the GPU.
    frames: CUDA (N,Hs,Ws) float32 mm or (N,Hs,Ws,3) uint8 BGR; jt_xyz (N,J,3) / center_xyz (N,3) float64 mm as nyu_loader.make_dataset holds
    them; cube (3,) (NYU.cube); augs: N tuples from random_aug.  Returns (img CUDA (N,1,D,D) f32, jt_xyz (N,J,3), jt_uvd (N,J,3),
    center_xyz (N,3), M (N,3,3), cube (N,3)) -- the label tensors are small float32 CPU tensors."""
    L.require_cuda(frames)
    if frames.dtype == torch.float32 and frames.dim() == 3:
        fmt = 0
    elif frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3:
        fmt = 1
    else:
        raise ValueError("frames must be (N,Hs,Ws) float32 or (N,Hs,Ws,3) uint8")
    frames = frames.contiguous()
    N, Hs, Ws = frames.shape[:3]
    if len(augs) != N:
        raise ValueError("one augmentation draw per frame")
    geo = train_batch_geometry(jt_xyz, center_xyz, cube, img_size, paras, flip, augs)
    params = torch.from_numpy(geo[0]).to(frames.device)
    out = torch.empty(N, 1, img_size, img_size, dtype=torch.float32, device=frames.device)
    L.check(L.lib().awr_crop_augment_normalize(L.ptr(frames), fmt, N, Hs, Ws, L.ptr(params), int(img_size), L.ptr(out), L.stream()),
            "awr_crop_augment_normalize")
    return (out,) + tuple(torch.from_numpy(np.ascontiguousarray(g)) for g in geo[1:])
