"""Small numpy geometry helpers for the synthetic-graph generators and host-side checks.

Conventions follow the reference's value types (paths relative to the reference checkout):
  * rotation  wRo = Rz(yaw) * Ry(pitch) * Rx(roll)   (pop_planar_slam/Thirdparty/isam/include/isam/Rot3d.h:55-82)
  * pose value = (x, y, z, qw, qx, qy, qz)           (Pose3d.h:78-79)
  * pose measurement = (x, y, z, yaw, pitch, roll)   (Pose3d::vector, Pose3d.h:138-145)
  * plane = unit homogeneous 4-vector (a, b, c, d)   (pop_planar_slam/src/isam_plane3d.h:27-66)
  * plane update: q' = Q(delta) (x) q_pi, q_pi = (w=d; x,y,z=a,b,c)   (isam_plane3d.h:101-127)

This module is host-side product code (generators for bench.py / tests); it never touches oracle/.
"""
import numpy as np


def euler_to_R(yaw, pitch, roll):
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


def R_to_euler(R):
    """(yaw, pitch, roll) of R = Rz Ry Rx; equals quat_to_euler(quat(R)) of Rot3d.h:114-124."""
    yaw = np.arctan2(R[1, 0], R[0, 0])
    pitch = np.arcsin(np.clip(-R[2, 0], -1.0, 1.0))
    roll = np.arctan2(R[2, 1], R[2, 2])
    return yaw, pitch, roll


def quat_to_R(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def R_to_quat(R):
    """(w,x,y,z), trace / largest-diagonal method."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([w, (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    v = np.zeros(3)
    v[i] = 0.5 * s
    s = 0.5 / s
    w = (R[k, j] - R[j, k]) * s
    v[j] = (R[j, i] + R[i, j]) * s
    v[k] = (R[k, i] + R[i, k]) * s
    return np.array([w, v[0], v[1], v[2]])


def pose7_to_T(p):
    T = np.eye(4)
    T[:3, :3] = quat_to_R(p[3:7])
    T[:3, 3] = p[:3]
    return T


def T_to_pose7(T):
    q = R_to_quat(T[:3, :3])
    return np.concatenate([T[:3, 3], q])


def T_to_xyzypr(T):
    y, p, r = R_to_euler(T[:3, :3])
    return np.array([T[0, 3], T[1, 3], T[2, 3], y, p, r])


def xyzypr_to_T(v):
    T = np.eye(4)
    T[:3, :3] = euler_to_R(v[3], v[4], v[5])
    T[:3, 3] = v[:3]
    return T


def inv_T(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
    ])


def delta3_to_quat(d):
    th = float(np.linalg.norm(d))
    if th < 1e-12:
        return np.array([1.0, 0.5 * d[0], 0.5 * d[1], 0.5 * d[2]])
    s = np.sin(0.5 * th) / th
    return np.array([np.cos(0.5 * th), s * d[0], s * d[1], s * d[2]])


def plane_normalize(p):
    p = np.asarray(p, dtype=np.float64)
    return p / np.linalg.norm(p)


def plane_exmap(p, d):
    """isam_plane3d.h:101-127 (plane_type == -1)."""
    qp = np.array([p[3], p[0], p[1], p[2]])
    q = quat_mul(delta3_to_quat(d), qp)
    return plane_normalize(np.array([q[1], q[2], q[3], q[0]]))


def plane_to_local(T_wo, p):
    """Plane3d::transform_to(wTo) = normalize(wTo^T pi)  (isam_plane3d.h:180-182)."""
    return plane_normalize(T_wo.T @ p)


def plane_to_global(T_wo, p_local):
    """Plane3d::transform_from(oTw) = normalize(oTw^T pi_local)  (isam_plane3d.h:186-188)."""
    return plane_normalize(inv_T(T_wo).T @ p_local)


def plane_distance(pa, pb):
    """Sign-invariant distance between two planes given as unit 4-vectors."""
    return min(np.linalg.norm(pa - pb), np.linalg.norm(pa + pb))


def quat_angle(qa, qb):
    """Rotation angle between two unit quaternions (sign invariant)."""
    d = abs(float(np.dot(qa, qb)))
    return 2.0 * np.arccos(min(1.0, d))
