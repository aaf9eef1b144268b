"""Quaternion and Euler-angle utilities (passive rotations, scalar-first quaternions).

Semantics follow the reference's helpers (machupX/helpers.py:151-248): ``earth_to_body(q, v)``
expresses an earth-frame vector in the frame whose orientation is q (reference ``quat_trans``);
``body_to_earth`` is its inverse (reference ``quat_inv_trans``).  Implemented here through the
direction-cosine matrix so whole (n, 3) tables rotate with one matmul.
"""
import numpy as np


def quat_normalize(q):
    q = np.asarray(q, dtype=float)
    return q / np.sqrt(np.dot(q, q))


def rotation_matrix(q):
    """3x3 matrix C with v_body = C @ v_earth for orientation quaternion q = [e0, ex, ey, ez]."""
    e0, ex, ey, ez = (float(c) for c in q)
    return np.array([
        [ex * ex + e0 * e0 - ey * ey - ez * ez, 2.0 * (ex * ey + ez * e0), 2.0 * (ex * ez - ey * e0)],
        [2.0 * (ex * ey - ez * e0), ey * ey + e0 * e0 - ex * ex - ez * ez, 2.0 * (ey * ez + ex * e0)],
        [2.0 * (ex * ez + ey * e0), 2.0 * (ey * ez - ex * e0), ez * ez + e0 * e0 - ex * ex - ey * ey],
    ])


def earth_to_body(q, v):
    v = np.asarray(v, dtype=float)
    return v @ rotation_matrix(q).T


def body_to_earth(q, v):
    v = np.asarray(v, dtype=float)
    return v @ rotation_matrix(q)


def quat_conjugate(q):
    return np.array([q[0], -q[1], -q[2], -q[3]], dtype=float)


def quat_product(a, b):
    """Hamilton product; 3-vectors are promoted to pure quaternions."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if a.shape[0] == 3:
        a = np.concatenate(([0.0], a))
    if b.shape[0] == 3:
        b = np.concatenate(([0.0], b))
    a0, av = a[0], a[1:]
    b0, bv = b[0], b[1:]
    out = np.empty(4)
    out[0] = a0 * b0 - np.dot(av, bv)
    out[1:] = a0 * bv + b0 * av + np.cross(av, bv)
    return out


def euler_to_quat(E):
    """[phi, theta, psi] (rad) -> quaternion (Phillips, Mechanics of Flight, eq. 11.7.8)."""
    hp, ht, hs = 0.5 * E[0], 0.5 * E[1], 0.5 * E[2]
    cp, sp = np.cos(hp), np.sin(hp)
    ct, st = np.cos(ht), np.sin(ht)
    cs, ss = np.cos(hs), np.sin(hs)
    return np.array([
        cp * ct * cs + sp * st * ss,
        sp * ct * cs - cp * st * ss,
        cp * st * cs + sp * ct * ss,
        cp * ct * ss - sp * st * cs,
    ])


def quat_to_euler(q):
    """Quaternion -> [phi, theta, psi] (rad) (Phillips eq. 11.7.11), gimbal-lock aware."""
    e0, ex, ey, ez = (float(c) for c in q)
    t = e0 * ey - ex * ez
    if t == 0.5 or t == -0.5:
        return np.array([2.0 * np.arcsin(ex / np.cos(np.pi / 4)), np.pi * t, 0.0])
    return np.array([
        np.arctan2(2.0 * (e0 * ex + ey * ez), e0 * e0 + ez * ez - ex * ex - ey * ey),
        np.arcsin(2.0 * t),
        np.arctan2(2.0 * (e0 * ez + ex * ey), e0 * e0 + ex * ex - ey * ey - ez * ez),
    ])
