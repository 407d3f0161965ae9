"""numpy restatement of the camodocal camera models' liftProjective / spaceToPlane (TEST INFRASTRUCTURE; only tests/ import it).
    PINHOLE          camera_model/src/camera_models/PinholeCamera.cc:450-510, 530-560
    MEI              camera_model/src/camera_models/CataCamera.cc:556-625 (lift), 632-658 (spaceToPlane), 766-782 (distortion)
    KANNALA_BRANDT   camera_model/src/camera_models/EquidistantCamera.cc:428-461, backprojectSymmetric :716-818
backprojectSymmetric takes the smallest non-negative real eigenvalue of the polynomial's companion matrix; numpy.roots computes
exactly those eigenvalues."""
import numpy as np


def _distortion(k1, k2, p1, p2, x, y):
    x2, y2, xy = x * x, y * y, x * y
    rho2 = x2 + y2
    rad = k1 * rho2 + k2 * rho2 * rho2
    return x * rad + 2.0 * p1 * xy + p2 * (rho2 + 2.0 * x2), y * rad + 2.0 * p2 * xy + p1 * (rho2 + 2.0 * y2)


def _undistort_recursive(k1, k2, p1, p2, mx_d, my_d, n=8):
    if k1 == 0.0 and k2 == 0.0 and p1 == 0.0 and p2 == 0.0:
        return mx_d, my_d
    dx, dy = _distortion(k1, k2, p1, p2, mx_d, my_d)
    mx_u, my_u = mx_d - dx, my_d - dy
    for _ in range(1, n):
        dx, dy = _distortion(k1, k2, p1, p2, mx_u, my_u)
        mx_u, my_u = mx_d - dx, my_d - dy
    return mx_u, my_u


def mei_lift(px, py, xi, k1, k2, p1, p2, gamma1, gamma2, u0, v0):
    mx_d = (1.0 / gamma1) * px + (-u0 / gamma1)
    my_d = (1.0 / gamma2) * py + (-v0 / gamma2)
    mx_u, my_u = _undistort_recursive(k1, k2, p1, p2, mx_d, my_d)
    if xi == 1.0:
        z = (1.0 - mx_u * mx_u - my_u * my_u) / 2.0
    else:
        rho2 = mx_u * mx_u + my_u * my_u
        z = 1.0 - xi * (rho2 + 1.0) / (xi + np.sqrt(1.0 + (1.0 - xi * xi) * rho2))
    return np.array([mx_u, my_u, z])


def mei_space_to_plane(P, xi, k1, k2, p1, p2, gamma1, gamma2, u0, v0):
    z = P[2] + xi * np.linalg.norm(P)
    x, y = P[0] / z, P[1] / z
    dx, dy = _distortion(k1, k2, p1, p2, x, y)
    return np.array([gamma1 * (x + dx) + u0, gamma2 * (y + dy) + v0])


def kb_backproject_symmetric(r, k2, k3, k4, k5, tol=1e-10):
    ks = [k2, k3, k4, k5]
    npow = 9
    for k in reversed(ks):
        if k == 0.0:
            npow -= 2
    # the reference shrinks the degree by 2 for EVERY zero coefficient, whichever it is (EquidistantCamera.cc:733-749)
    coeffs = np.zeros(npow + 1)
    coeffs[0] = -r
    coeffs[1] = 1.0
    for i, k in zip((3, 5, 7, 9), ks):
        if npow >= i:
            coeffs[i] = k
    if npow == 1:
        return r
    roots = np.roots(coeffs[::-1])
    thetas = []
    for z in roots:
        if abs(z.imag) > tol:
            continue
        t = z.real
        if t < -tol:
            continue
        thetas.append(max(t, 0.0))
    return min(thetas) if thetas else r


def kb_lift(px, py, k2, k3, k4, k5, mu, mv, u0, v0):
    x = (1.0 / mu) * px + (-u0 / mu)
    y = (1.0 / mv) * py + (-v0 / mv)
    r = np.hypot(x, y)
    phi = 0.0 if r < 1e-10 else np.arctan2(y, x)
    theta = kb_backproject_symmetric(r, k2, k3, k4, k5)
    return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])


def kb_space_to_plane(P, k2, k3, k4, k5, mu, mv, u0, v0):
    theta = np.arccos(P[2] / np.linalg.norm(P))
    phi = np.arctan2(P[1], P[0])
    t2 = theta * theta
    rr = theta * (1 + t2 * (k2 + t2 * (k3 + t2 * (k4 + t2 * k5))))
    return np.array([mu * rr * np.cos(phi) + u0, mv * rr * np.sin(phi) + v0])
