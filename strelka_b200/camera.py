"""oka::Camera mirror (reference: include/scene/camera.h:16-95, src/scene/camera.cpp).

Only what the render path consumes: position + orientation quaternion -> view matrix
(camera.cpp:10-23, first-person: view = R(q) * T(-pos)), fov (degrees).  The projection is
re-derived by the backend from the output buffer's aspect ratio (OptixRender.cpp:895-897).
"""
from __future__ import annotations

import numpy as np


def _quat_to_mat(q):
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 0.0],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 0.0],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), 0.0],
            [0.0, 0.0, 0.0, 1.0],
        ],
        dtype=np.float64,
    )


def _mat_to_quat(m):
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        return np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    i = int(np.argmax([m[0, 0], m[1, 1], m[2, 2]]))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
    q = np.zeros(4)
    q[0] = (m[k, j] - m[j, k]) / s
    q[1 + i] = 0.25 * s
    q[1 + j] = (m[j, i] + m[i, j]) / s
    q[1 + k] = (m[k, i] + m[i, k]) / s
    return q


class Camera:
    def __init__(self):
        self.position = np.zeros(3, dtype=np.float64)
        self.orientation = np.array([1.0, 0.0, 0.0, 0.0])  # w, x, y, z (glm::quat mOrientation)
        self.fov = 45.0
        self.znear = 0.1
        self.zfar = 1000.0
        self.view = np.eye(4, dtype=np.float64)  # math convention: p_view = view @ p_world
        self.updateViewMatrix()

    # camera.cpp:10-23 (CameraType::firstperson)
    def updateViewMatrix(self):  # noqa: N802
        rot = _quat_to_mat(self.orientation)
        trans = np.eye(4)
        trans[:3, 3] = -self.position
        self.view = rot @ trans

    def setPosition(self, p):  # noqa: N802
        self.position = np.asarray(p, dtype=np.float64)
        self.updateViewMatrix()

    def setFov(self, fov):  # noqa: N802
        self.fov = float(fov)

    def look_at(self, eye, target, up=(0.0, 1.0, 0.0)):
        """Convenience: orientation such that the camera at `eye` looks at `target` (-Z forward)."""
        eye = np.asarray(eye, dtype=np.float64)
        f = np.asarray(target, dtype=np.float64) - eye
        f /= np.linalg.norm(f)
        r = np.cross(f, np.asarray(up, dtype=np.float64))
        r /= np.linalg.norm(r)
        u = np.cross(r, f)
        rot = np.eye(4)
        rot[0, :3], rot[1, :3], rot[2, :3] = r, u, -f  # world -> view rotation
        self.orientation = _mat_to_quat(rot[:3, :3])
        self.position = eye
        self.updateViewMatrix()

    def view_glm(self) -> np.ndarray:
        """The view matrix in glm storage order (column-major float32[16]) for sb_set_camera."""
        return np.ascontiguousarray(self.view.T.astype(np.float32).reshape(16))
