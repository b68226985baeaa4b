"""Independent float64 numpy restatement of the reference's render-and-compare backward as a TENSOR program
(python/stillleben/diff.py:73-127, 355-523): masks via the mask oracle, then whole-image einsum contractions in
the order of the reference's bmm chain. Used to pin oracle/orc_diff.cpp:orc_diff_pose_grad (a per-pixel loop)."""
import ctypes as C

import numpy as np

import oracle_util as ou


def masks(inst, depth):
    H, W = inst.shape
    valid = np.zeros((H, W), np.uint8)
    ou.lib().orc_diff_sobel_valid_mask(inst.ctypes.data, depth.ctypes.data, valid.ctypes.data, H, W)
    return valid


def dilate(mask, valid, coord4):
    H, W = mask.shape
    mo = np.zeros((H, W), np.uint8)
    co = np.zeros((H, W, 3), np.float32)
    ou.lib().orc_diff_dilate_object_mask(mask.ctypes.data, valid.ctypes.data, coord4.ctypes.data, 4, mo.ctypes.data, co.ctypes.data, H, W)
    return mo.astype(bool), co


def pose_grad(rgb, inst, coord4, grad_img, P, poses, ids):
    """rgb HxWx4 u8, inst HxW i16, coord4 HxWx4 f32, grad_img 3xHxW, P 4x4 and poses Nx4x4 indexed [row, col]."""
    H, W = inst.shape
    depth = np.ascontiguousarray(coord4[..., 3])
    valid = masks(inst, depth)
    img = rgb[..., :3].astype(np.float64).transpose(2, 0, 1) / 255.0
    px = np.pad(img, ((0, 0), (0, 0), (1, 1)))
    py = np.pad(img, ((0, 0), (1, 1), (0, 0)))
    gx = -(px[:, :, 2:] - px[:, :, :-2]) / (2.0 / W * 2.0)
    gy = -(py[:, 2:, :] - py[:, :-2, :]) / (2.0 / H * 2.0)
    gx[:, valid == 0] = 0
    gy[:, valid == 0] = 0
    g_xy_img = np.stack([gx, gy], 0)                                   # 2 x 3 x H x W
    gens = np.zeros((6, 4, 4))
    gens[0, 1, 2], gens[0, 2, 1] = -1, 1
    gens[1, 0, 2], gens[1, 2, 0] = 1, -1
    gens[2, 0, 1], gens[2, 1, 0] = -1, 1
    gens[3, 0, 3] = gens[4, 1, 3] = gens[5, 2, 3] = 1
    out = np.zeros((len(ids), 6))
    P = np.asarray(P, np.float64)
    for o, idx in enumerate(ids):
        T0 = np.asarray(poses[o], np.float64)
        m, oc = dilate((inst == idx).astype(np.uint8), valid, coord4)
        if not m.any():
            continue
        x = np.concatenate([oc[m].astype(np.float64).T, np.ones((1, m.sum()))], 0)      # 4 x N
        y = T0 @ x
        den = P[2:3] @ y
        g_coord = np.zeros((2, 3, x.shape[1]))
        for j in range(2):
            for i in range(3):
                g_coord[j, i] = P[j, i] * (1 / den) + (P[2, i] * (-1 / den ** 2)) * (P[j:j + 1] @ y)
        g_pose = np.stack([(T0 @ gens[k] @ x)[:3] for k in range(6)], 0)               # 6 x 3 x N
        g_xy = g_xy_img[:, :, m]                                                        # 2 x 3 x N
        A = np.einsum("jcn,jin->nci", g_xy, g_coord)                                    # [N x 3 x 2] @ [N x 2 x 3]
        B = np.einsum("nci,kin->nck", A, g_pose)                                        # @ [N x 3 x 6]
        g_in = grad_img[:, m].astype(np.float64)                                        # 3 x N
        out[o] = np.einsum("cn,nck->k", g_in, B)
    return out


def oracle_pose_grad(rgb, inst, coord4, grad_img, P, poses, ids):
    H, W = inst.shape
    out = np.zeros((len(ids), 6), np.float32)
    Pr = np.ascontiguousarray(P, np.float32)
    Tr = np.ascontiguousarray(poses, np.float32)
    idv = np.ascontiguousarray(ids, np.int32)
    g = np.ascontiguousarray(grad_img, np.float32)
    L = ou.lib()
    L.orc_diff_pose_grad.restype = None
    L.orc_diff_pose_grad.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.orc_diff_pose_grad(rgb.ctypes.data, inst.ctypes.data, coord4.ctypes.data, g.ctypes.data, Pr.ctypes.data, Tr.ctypes.data,
                         idv.ctypes.data, len(ids), out.ctypes.data, H, W)
    return out


def synthetic_inputs(seed=0, H=60, W=80, n_obj=4):
    """Blobby instance map with touching / occluding objects, plausible object coordinates, random rgb and dL/dI."""
    rng = np.random.RandomState(seed)
    inst = np.zeros((H, W), np.int16)
    depth = np.full((H, W), 3000.0, np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    ids = []
    for k in range(n_obj):
        cy, cx, r = rng.uniform(10, H - 10), rng.uniform(10, W - 10), rng.uniform(6, 16)
        d = rng.uniform(0.5, 2.0)
        m = ((yy - cy) ** 2 + (xx - cx) ** 2 < r * r) & (d < depth)
        inst[m] = k + 1
        depth[m] = d + 0.01 * rng.rand(int(m.sum()))
        ids.append(k + 1)
    coord4 = np.zeros((H, W, 4), np.float32)
    coord4[..., :3] = rng.uniform(-0.2, 0.2, size=(H, W, 3))
    coord4[..., :3][inst == 0] = 3000.0
    coord4[..., 3] = depth
    rgb = rng.randint(0, 256, size=(H, W, 4)).astype(np.uint8)
    grad = rng.normal(size=(3, H, W)).astype(np.float32)
    P = np.array([[3.3, 0, 0.02, 0], [0, 4.4, -0.01, 0], [0, 0, 1.02, -0.2], [0, 0, 1, 0]], np.float32)
    poses = []
    for k in range(n_obj):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = rng.uniform(-0.3, 0.3, 3) + np.array([0, 0, 1.0])
        poses.append(T)
    ids.append(99)                     # an object that is not visible: its row stays zero
    poses.append(np.eye(4))
    return rgb, inst, coord4, grad, P, np.array(poses, np.float32), np.array(ids, np.int32)
