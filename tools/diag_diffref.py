import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref, diff_ref
from stillleben_b200 import sl, diff
ext = build_ref.load_diff()
sl.init_cuda(0)
dev = torch.device("cuda", 0)
for seed, H, W in ((1, 480, 640), (3, 480, 640), (0, 64, 96)):
    rgb, inst, coord4, grad, P, poses, ids = diff_ref.synthetic_inputs(seed, H=H, W=W)
    inst_t, depth_t = torch.from_numpy(inst).to(dev), torch.from_numpy(np.ascontiguousarray(coord4[..., 3])).to(dev)
    coord_t = torch.from_numpy(np.ascontiguousarray(coord4[..., :3])).to(dev)
    valid_ref = ext.generate_sobel_valid_mask(inst_t, depth_t)
    valid_mine = diff.generate_sobel_valid_mask(inst_t, depth_t)
    valid_orc = diff_ref.masks(inst, np.ascontiguousarray(coord4[..., 3])).astype(bool)
    print(seed, H, W, "valid ref==mine", torch.equal(valid_ref, valid_mine), "ref==orc", np.array_equal(valid_ref.cpu().numpy(), valid_orc), "n_invalid", int((~valid_orc).sum()))
    for idx in ids[:-1]:
        mask_t = inst_t == int(idx)
        m_ref, c_ref = ext.dilate_object_mask(mask_t, valid_ref, coord_t)
        m_mine, c_mine = diff.dilate_object_mask(mask_t, valid_mine, coord_t)
        m_orc, c_orc = diff_ref.dilate((inst == idx).astype(np.uint8), valid_orc.astype(np.uint8), coord4)
        bad = (c_ref != c_mine).any(-1)
        print("  obj", idx, "mask ref==mine", torch.equal(m_ref, m_mine), "ref==orc", np.array_equal(m_ref.cpu().numpy(), m_orc),
              "coords ref==mine", torch.equal(c_ref, c_mine), "ref==orc", np.array_equal(c_ref.cpu().numpy(), c_orc), "nbad", int(bad.sum()))
        if bad.any():
            ys, xs = torch.nonzero(bad, as_tuple=True)
            for y, x in list(zip(ys.tolist(), xs.tolist()))[:5]:
                print("    at", y, x, "ref", c_ref[y, x].tolist(), "mine", c_mine[y, x].tolist(), "mask", int(mask_t[y, x]), "m_ref", int(m_ref[y, x]),
                      "nbr mask", mask_t[max(y-1,0):y+2, max(x-1,0):x+2].int().tolist(), "valid", valid_ref[max(y-1,0):y+2, max(x-1,0):x+2].int().tolist())
