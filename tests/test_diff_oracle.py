"""CPU pins of the render-and-compare backward oracle (oracle/orc_diff.cpp:orc_diff_pose_grad): the per-pixel
restatement must agree with an independent float64 tensor-program restatement of diff.py (tests/diff_ref.py),
and with central finite differences of the quantity the chain rule differentiates."""
import numpy as np

import diff_ref


def test_oracle_matches_tensor_program():
    for seed in (0, 1, 2):
        args = diff_ref.synthetic_inputs(seed)
        ref = diff_ref.pose_grad(*args)
        got = diff_ref.oracle_pose_grad(*args)
        assert np.abs(ref[:-1]).max() > 1.0                      # the case is not degenerate
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-4 * np.abs(ref).max())
        assert (got[-1] == 0).all()                              # invisible object: zero row (diff.py:411-415)


def test_generators_are_the_derivative_of_the_linearised_pose():
    """g_coord @ g_pose is d(projected xy)/d(alpha..c) of T0 * delta(alpha..c) * x with the reference's
    row-2 divisor: check the closed form the oracle uses against finite differences."""
    rng = np.random.RandomState(3)
    P = np.array([[3.3, 0, 0.02, 0], [0, 4.4, -0.01, 0], [0, 0, 1.02, -0.2], [0, 0, 1, 0]])
    T0 = np.eye(4); T0[:3, 3] = (0.1, -0.2, 1.3)
    x = np.array([0.05, -0.03, 0.08, 1.0])

    def proj(delta):
        al, be, ga, a, b, c = delta
        D = np.array([[1, -ga, be, a], [ga, 1, -al, b], [-be, al, 1, c], [0, 0, 0, 1.0]])
        y = T0 @ D @ x
        return (P[:2] @ y) / (P[2] @ y)

    eps = 1e-6
    J = np.stack([(proj(np.eye(6)[k] * eps) - proj(-np.eye(6)[k] * eps)) / (2 * eps) for k in range(6)], 1)   # 2 x 6
    # one-pixel image: instance 1 everywhere valid, gradient picks dI/dx = 1 on channel 0 through rgb differences
    H = W = 3
    inst = np.ones((H, W), np.int16)
    coord4 = np.zeros((H, W, 4), np.float32); coord4[..., :3] = x[:3]; coord4[..., 3] = 1.0
    rgb = np.zeros((H, W, 4), np.uint8); rgb[:, 2, 0] = 255            # centre pixel: (right - left)/255 = 1
    grad = np.zeros((3, H, W), np.float32); grad[0, 1, 1] = 1.0         # only the centre pixel carries dL/dI
    out = diff_ref.oracle_pose_grad(rgb, inst, coord4, grad, P.astype(np.float32), T0[None].astype(np.float32), np.array([1], np.int32))
    gx_centre = -(1.0 - 0.0) * W / 4.0                                   # the reference's scaled, negated central difference
    np.testing.assert_allclose(out[0], gx_centre * J[0], rtol=1e-3, atol=1e-5)
    del rng
