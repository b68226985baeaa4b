"""Comparison of a CUDA frame with an oracle frame under the tolerances of BASELINE.json's north_star:
ID maps bit-exact; float targets within 1e-3 relative (plus an absolute floor for values near zero);
the RGBA8 colour target within 1 LSB.  Returns per-target statistics; `assert_parity` enforces them."""
import numpy as np

RTOL = 1e-3
ATOL = {"coord": 2e-5, "normals": 2e-5, "barycentric": 2e-5, "cam_coord": 2e-5, "hdr": 1e-4}
EXACT = ("class_index", "instance_index", "vertex_index")


def compare(gpu, ref):
    stats = {}
    for name, g in gpu.items():
        r = ref[name]
        assert g.shape == r.shape, (name, g.shape, r.shape)
        if name in EXACT:
            stats[name] = {"mismatch": int((g != r).sum()), "n": int(g.size)}
        elif name == "rgb":
            d = np.abs(g.astype(np.int32) - r.astype(np.int32))
            stats[name] = {"max": int(d.max()), "over1": int((d > 1).any(axis=-1).sum()), "n": int(d.shape[0] * d.shape[1])}
        else:
            g64, r64 = g.astype(np.float64), r.astype(np.float64)
            both_nan = np.isnan(g64) & np.isnan(r64)
            err = np.abs(g64 - r64)
            tol = RTOL * np.abs(r64) + ATOL.get(name, 2e-5)
            bad = (err > tol) & ~both_nan
            bad |= np.isnan(g64) != np.isnan(r64)
            stats[name] = {"max_abs": float(np.nanmax(err)) if err.size else 0.0, "bad": int(bad.any(axis=-1).sum()),
                           "n": int(g.shape[0] * g.shape[1])}
    return stats


def assert_parity(gpu, ref, rgb_outlier_frac=0.0, hdr_outlier_frac=0.0, rgb_outliers=None):
    """rgb_outliers: absolute number of RGBA8 pixels allowed beyond 1 LSB (overrides rgb_outlier_frac); hdr_outlier_frac:
    share of HDR pixels allowed beyond RTOL (at least 4 pixels when non-zero)."""
    st = compare(gpu, ref)
    for name, s in st.items():
        if name in EXACT:
            assert s["mismatch"] == 0, (name, s)
        elif name == "rgb":
            assert s["over1"] <= (rgb_outliers if rgb_outliers is not None else rgb_outlier_frac * s["n"]), (name, s)
        elif name == "hdr":
            assert s["bad"] <= (max(4, hdr_outlier_frac * s["n"]) if hdr_outlier_frac else 0), (name, s)
        else:
            assert s["bad"] == 0, (name, s)
    return st
