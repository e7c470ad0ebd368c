"""Oracle of the landmark bundle adjustment -- TEST INFRASTRUCTURE, not product code.

What the reference builds with DO_FEATURE_OPTIMIZATION (src/landmark.cpp:97-187, src/graph_manager.cpp:963-967):
  * VertexSE3 camera poses (world-from-camera, X <- X * fromVectorMQT(delta) as in g2o),
  * VertexPointXYZ landmarks,
  * one EdgeSE3PointXYZDepth per observation: error = (fx x/z + cx - u, fy y/z + cy - v, z - depth) of the landmark in the camera
    frame, information diag(1, 1, 1 / depth_covariance(depth)) (misc2.h:37-47), camera ParameterCamera(fx, fy, cx, cy),
  * the camera-camera EdgeSE3 constraints (error toVectorMQT(Z^-1 Xi^-1 Xj), information matrix of the edge).
g2o is not under /root/reference: parity unpinned.  The reference does NOT marginalise the landmarks; this oracle solves the FULL
(6 Ncam + 3 Npoint) damped normal equations densely with numpy (Jacobians by central differences on the manifold), the CUDA path
eliminates the points by the Schur complement -- different algorithms, same optimum.
"""
from __future__ import annotations

import numpy as np


def quat_to_rot(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def pose_oplus(p, d):
    """VertexSE3::oplus: X <- X * fromVectorMQT(d), d = (t, qx, qy, qz), w = sqrt(1 - |q|^2) (identity rotation if |q| > 1)."""
    R = quat_to_rot(p[3:])
    t = p[:3] + R @ d[:3]
    w2 = 1.0 - float(d[3:] @ d[3:])
    if w2 < 0:
        return np.concatenate([t, p[3:]])
    q = quat_mul(p[3:], np.array([d[3], d[4], d[5], np.sqrt(w2)]))
    return np.concatenate([t, q / np.linalg.norm(q)])


def obs_error(pose, pw, uvd, K4):
    """EdgeSE3PointXYZDepth::computeError"""
    R = quat_to_rot(pose[3:])
    pc = R.T @ (pw - pose[:3])
    return np.array([K4[0] * pc[0] / pc[2] + K4[2] - uvd[0], K4[1] * pc[1] / pc[2] + K4[3] - uvd[1], pc[2] - uvd[2]])


def edge_error(pi, pj, z):
    """EdgeSE3::computeError: toVectorMQT(Z^-1 Xi^-1 Xj)"""
    Ri, Rj, Rz = quat_to_rot(pi[3:]), quat_to_rot(pj[3:]), quat_to_rot(z[3:])
    tb = Ri.T @ (pj[:3] - pi[:3])
    et = Rz.T @ (tb - z[:3])
    conj = lambda q: np.array([-q[0], -q[1], -q[2], q[3]])
    nq = lambda q: q / np.linalg.norm(q)
    qe = quat_mul(quat_mul(conj(nq(z[3:])), conj(nq(pi[3:]))), nq(pj[3:]))
    qe = nq(qe)
    if qe[3] < 0:
        qe = -qe
    return np.concatenate([et, qe[:3]])


class Problem:
    def __init__(self, poses, fixed, points, obs_cam, obs_point, obs_uvd, obs_info3, K4, ij=None, meas=None, info=None, huber_delta=1.0):
        self.poses = np.array(poses, np.float64)
        self.fixed = np.asarray(fixed, bool)
        self.points = np.array(points, np.float64)
        self.oc, self.op = np.asarray(obs_cam, int), np.asarray(obs_point, int)
        self.uvd, self.w3 = np.asarray(obs_uvd, np.float64), np.asarray(obs_info3, np.float64)
        self.K4 = np.asarray(K4, np.float64)
        self.ij = np.zeros((0, 2), int) if ij is None else np.asarray(ij, int).reshape(-1, 2)
        self.meas = np.zeros((0, 7)) if meas is None else np.asarray(meas, np.float64).reshape(-1, 7)
        self.info = np.zeros((0, 36)) if info is None else np.asarray(info, np.float64).reshape(-1, 36)
        self.delta = huber_delta

    def residual_blocks(self, poses, points):
        """list of (variable ids, residual vector, information matrix, robust) -- variable id: ('c', k) or ('p', k)"""
        out = []
        for o in range(len(self.oc)):
            c, p = self.oc[o], self.op[o]
            out.append(((("c", c), ("p", p)), obs_error(poses[c], points[p], self.uvd[o], self.K4), np.diag(self.w3[o]), False))
        for k, (i, j) in enumerate(self.ij):
            out.append(((("c", i), ("c", j)), edge_error(poses[i], poses[j], self.meas[k]), self.info[k].reshape(6, 6), True))
        return out

    def chi2(self, poses=None, points=None, robust=True):
        poses = self.poses if poses is None else poses
        points = self.points if points is None else points
        tot = 0.0
        for _, e, W, rob in self.residual_blocks(poses, points):
            c = float(e @ W @ e)
            if rob and robust and c > self.delta ** 2:
                c = 2 * np.sqrt(c) * self.delta - self.delta ** 2
            tot += c
        return tot

    def _index(self):
        nc, npt = len(self.poses), len(self.points)
        return nc, npt, 6 * nc + 3 * npt

    def _apply(self, dx):
        nc, npt, _ = self._index()
        poses = np.stack([self.poses[c] if self.fixed[c] else pose_oplus(self.poses[c], dx[6 * c:6 * c + 6]) for c in range(nc)])
        points = self.points + dx[6 * nc:].reshape(npt, 3)
        return poses, points

    def normal_equations(self, eps=1e-6):
        nc, npt, n = self._index()
        H = np.zeros((n, n)); b = np.zeros(n)

        def sl(v):
            kind, k = v
            return slice(6 * k, 6 * k + 6) if kind == "c" else slice(6 * nc + 3 * k, 6 * nc + 3 * k + 3)

        def perturbed(v, d):
            kind, k = v
            if kind == "c":
                return ("c", k, pose_oplus(self.poses[k], d))
            return ("p", k, self.points[k] + d)

        def eval_block(bi, subst):
            poses, points = self.poses, self.points
            if subst is not None:
                kind, k, val = subst
                if kind == "c":
                    poses = poses.copy(); poses[k] = val
                else:
                    points = points.copy(); points[k] = val
            vs = blocks[bi][0]
            if vs[1][0] == "p":
                o = bi
                return obs_error(poses[self.oc[o]], points[self.op[o]], self.uvd[o], self.K4)
            e = bi - len(self.oc)
            i, j = self.ij[e]
            return edge_error(poses[i], poses[j], self.meas[e])

        blocks = self.residual_blocks(self.poses, self.points)
        for bi, (vs, e, W, rob) in enumerate(blocks):
            rho1 = 1.0
            if rob:
                c = float(e @ W @ e)
                if c > self.delta ** 2:
                    rho1 = self.delta / np.sqrt(c)  # Huber: rho'(e2)
            Js = []
            for v in vs:
                dim = 6 if v[0] == "c" else 3
                J = np.zeros((len(e), dim))
                for a in range(dim):
                    d = np.zeros(dim); d[a] = eps
                    J[:, a] = (eval_block(bi, perturbed(v, d)) - eval_block(bi, perturbed(v, -d))) / (2 * eps)
                if v[0] == "c" and self.fixed[v[1]]:
                    J[:] = 0
                Js.append(J)
            for a, va in enumerate(vs):
                b[sl(va)] += rho1 * Js[a].T @ W @ e
                for c2, vb in enumerate(vs):
                    H[sl(va), sl(vb)] += rho1 * Js[a].T @ W @ Js[c2]
        return H, b

    def optimize(self, iterations=10):
        """Levenberg-Marquardt with g2o's bookkeeping (lambda0 = 1e-5 max diag H, gain ratio with the +1e-3 guard, <= 10 trials)."""
        nc, npt, n = self._index()
        free = np.ones(n, bool)
        for c in range(nc):
            if self.fixed[c]:
                free[6 * c:6 * c + 6] = False
        lam, ni = None, 2.0
        for it in range(iterations):
            cur = self.chi2()
            H, b = self.normal_equations()
            if lam is None:
                lam = 1e-5 * np.max(np.diag(H)[free])
            ok = False
            for trial in range(10):
                A = H[np.ix_(free, free)] + lam * np.eye(int(free.sum()))
                dx = np.zeros(n)
                dx[free] = np.linalg.solve(A, -b[free])
                poses, points = self._apply(dx)
                new = self.chi2(poses, points)
                scale = float(dx @ (lam * dx - b)) + 1e-3
                rho = (cur - new) / scale
                if rho > 0 and np.isfinite(new):
                    alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                    lam *= max(1.0 / 3.0, alpha)
                    ni = 2.0
                    self.poses, self.points = poses, points
                    ok = True
                    break
                lam *= ni
                ni *= 2
            if not ok:
                break
        return self.chi2()
