"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the deterministic half of the reference's feeder augmentation.

``random_move`` (feeders/tools.py:58-101) draws an angle, a scale and a translation at ``move_time + 1`` node frames with
``np.random.choice`` and interpolates them per frame with ``np.linspace``; the x / y channels of every frame are then
rotated, scaled and translated.  ``move_nodes`` restates the draws (same calls, same order), ``apply_move`` the
arithmetic for given node values -- the part the CUDA kernel ``sgcn_random_move`` replaces.  Both are pinned against
the reference's own function, executed from /root/reference, by tests/golden/feeder.npz (oracle/make_golden.py step 9).
"""
import numpy as np

ANGLES = [-10., -5., 0., 5., 10.]
SCALES = [0.9, 1.0, 1.1]
SHIFTS = [-0.2, -0.1, 0.0, 0.1, 0.2]


def move_nodes(T, move_time=1):
    """feeders/tools.py:65-73 -- (node [K+1] int, vals [4, K+1] float64) for ONE sample; consumes np.random like the
    reference (four np.random.choice calls of num_node draws each, in the order A, S, T_x, T_y)"""
    node = np.arange(0, T, T * 1.0 / move_time).round().astype(int)
    node = np.append(node, T)
    num_node = len(node)
    A = np.random.choice(ANGLES, num_node)
    S = np.random.choice(SCALES, num_node)
    T_x = np.random.choice(SHIFTS, num_node)
    T_y = np.random.choice(SHIFTS, num_node)
    return node, np.stack([A, S, T_x, T_y]).astype(np.float64)


def apply_move(data_numpy, node, vals):
    """feeders/tools.py:75-101 for given node values; data (C, T, V, M) -> new array of the same dtype"""
    C, T, V, M = data_numpy.shape
    out = data_numpy.copy()
    a, s, t_x, t_y = np.zeros(T), np.zeros(T), np.zeros(T), np.zeros(T)
    for i in range(len(node) - 1):
        n = node[i + 1] - node[i]
        a[node[i]:node[i + 1]] = np.linspace(vals[0, i], vals[0, i + 1], n) * np.pi / 180
        s[node[i]:node[i + 1]] = np.linspace(vals[1, i], vals[1, i + 1], n)
        t_x[node[i]:node[i + 1]] = np.linspace(vals[2, i], vals[2, i + 1], n)
        t_y[node[i]:node[i + 1]] = np.linspace(vals[3, i], vals[3, i + 1], n)
    theta = np.array([[np.cos(a) * s, -np.sin(a) * s], [np.sin(a) * s, np.cos(a) * s]])
    for i_frame in range(T):
        xy = out[0:2, i_frame, :, :]
        new_xy = np.dot(theta[:, :, i_frame], xy.reshape(2, -1))
        new_xy[0] += t_x[i_frame]
        new_xy[1] += t_y[i_frame]
        out[0:2, i_frame, :, :] = new_xy.reshape(2, V, M)
    return out
