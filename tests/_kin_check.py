"""Independent numpy forward kinematics / momentum bookkeeping used by the physics known-answer tests."""
import numpy as np


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rodrigues(a, th):
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def momentum_energy(model, st, q, qd, gravity=0.0):
    """Total linear momentum, angular momentum about the world origin, kinetic+potential energy."""
    D = model.chain_len
    R = [quat_to_mat(st[3:7])]
    o = [st[0:3].copy()]
    w = [st[10:13].copy()]
    v = [st[7:10].copy()]
    for b in range(1, model.nb):
        par = 0 if (b - 1) % D == 0 else b - 1
        B = model.body[b].astype(np.float64)
        ob = o[par] + R[par] @ B[0:3]
        a = R[par] @ B[3:6]
        o.append(ob)
        R.append(R[par] @ rodrigues(B[3:6], q[b - 1]))
        w.append(w[par] + a * qd[b - 1])
        v.append(v[par] + np.cross(w[par], ob - o[par]))  # velocity of the joint origin (fixed in parent)
    P, L, E = np.zeros(3), np.zeros(3), 0.0
    for b in range(model.nb):
        B = model.body[b].astype(np.float64)
        m = B[9]
        c = o[b] + R[b] @ B[6:9]
        vc = v[b] + np.cross(w[b], c - o[b])
        Il = np.array([[B[10], B[13], B[14]], [B[13], B[11], B[15]], [B[14], B[15], B[12]]])
        Iw = R[b] @ Il @ R[b].T
        P += m * vc
        L += Iw @ w[b] + m * np.cross(c, vc)
        E += 0.5 * m * vc @ vc + 0.5 * w[b] @ Iw @ w[b] + m * gravity * c[2]
    return P, L, E
