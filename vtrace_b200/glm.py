"""Host-side matrix builders the Rust engine gets from glm-rs 0.2.3 (Cargo.lock:182-185).

Call sites restated: src/render.rs:190,208-214 (perspective, look_at), src/render.rs:54-78
(GPUInstance::new / from_model), src/world.rs:151,181,189 (translate / scale).  glm-rs is not
vendored in the reference tree, so these are the standard right-handed GL formulas it
documents; parity for the traversal is defined on identical matrix *inputs*, so nothing
downstream depends on how these bits were produced ("parity unpinned" for this file).

All matrices are float32, column-major: m[col][row], exactly the 64-byte layout that crosses
the C ABI (lib/memory.c:392-406).
"""
from __future__ import annotations

import numpy as np

F = np.float32

# src/render.rs:190 — note the truncated pi literal
REFERENCE_FOV = F(80.0) / F(180.0) * F(3.1415926)
REFERENCE_NEAR = F(0.01)     # src/render.rs:209
REFERENCE_FAR = F(10000.0)   # src/render.rs:209


def identity() -> np.ndarray:
    return np.eye(4, dtype=F)


def perspective(fovy, aspect, near, far) -> np.ndarray:
    """glm::ext::perspective — RH, depth -1..1, no Y flip."""
    fovy, aspect, near, far = F(fovy), F(aspect), F(near), F(far)
    ys = F(1.0) / F(np.tan(fovy / F(2.0)))
    xs = ys / aspect
    m = np.zeros((4, 4), dtype=F)
    m[0][0] = xs
    m[1][1] = ys
    m[2][2] = (far + near) / (near - far)
    m[2][3] = F(-1.0)
    m[3][2] = (F(2.0) * far * near) / (near - far)
    return m


def _normalize(v):
    v = np.asarray(v, dtype=F)
    return v / F(np.sqrt(F(np.dot(v, v))))


def look_at(eye, center, up) -> np.ndarray:
    """glm::ext::look_at (RH)."""
    eye = np.asarray(eye, dtype=F)
    f = _normalize(np.asarray(center, dtype=F) - eye)
    s = _normalize(np.cross(f, np.asarray(up, dtype=F)).astype(F))
    u = np.cross(s, f).astype(F)
    m = identity()
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0] = -F(np.dot(s, eye))
    m[3][1] = -F(np.dot(u, eye))
    m[3][2] = F(np.dot(f, eye))
    return m


def translate(m, v) -> np.ndarray:
    m = np.array(m, dtype=F)
    v = np.asarray(v, dtype=F)
    r = m.copy()
    r[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3]
    return r


def scale(m, v) -> np.ndarray:
    m = np.array(m, dtype=F)
    v = np.asarray(v, dtype=F)
    r = m.copy()
    r[0], r[1], r[2] = m[0] * v[0], m[1] * v[1], m[2] * v[2]
    return r


def rotate(m, angle, axis) -> np.ndarray:
    m = np.array(m, dtype=F)
    a = F(angle)
    c, s = F(np.cos(a)), F(np.sin(a))
    ax = _normalize(axis)
    t = ax * (F(1.0) - c)
    rot = np.zeros((3, 3), dtype=F)
    rot[0][0] = c + t[0] * ax[0]
    rot[0][1] = t[0] * ax[1] + s * ax[2]
    rot[0][2] = t[0] * ax[2] - s * ax[1]
    rot[1][0] = t[1] * ax[0] - s * ax[2]
    rot[1][1] = c + t[1] * ax[1]
    rot[1][2] = t[1] * ax[2] + s * ax[0]
    rot[2][0] = t[2] * ax[0] + s * ax[1]
    rot[2][1] = t[2] * ax[1] - s * ax[0]
    rot[2][2] = c + t[2] * ax[2]
    r = m.copy()
    for j in range(3):
        r[j] = m[0] * rot[j][0] + m[1] * rot[j][1] + m[2] * rot[j][2]
    return r


def with_texture_id(model, texture_id: int) -> np.ndarray:
    """GPUInstance::from_model (src/render.rs:74-78): id bit-cast into element [3][3]."""
    r = np.array(model, dtype=F)
    r.reshape(16).view(np.uint32)[15] = np.uint32(texture_id)
    return r
