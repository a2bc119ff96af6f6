"""float32 vector / matrix helpers with the rounding behaviour of polaris' Go `types` package.

The tracer only ever receives `camera.Position` and `camera.Frustrum` (reference
tracer/opencl/tracer.go:176-179) and instance matrices that were inverted by
`Mat4.Inv()` (asset/compiler/compiler.go:191).  To hand the CUDA backend the same numbers
the Go host would, these helpers evaluate the same expressions in the same order in
IEEE float32 (numpy scalar arithmetic never widens float32 x float32):

  * Mat4 is column-major, 16 floats (types/matrix.go:61-69);
  * Mul4 / Mul4x1 sum k = 0..3 left to right (matrix.go:61-90);
  * Inv is the cofactor expansion of matrix.go:108-138, term order preserved, returning
    the zero matrix when |det| < 1e-10;
  * Perspective4 treats fovy as RADIANS (the deg->rad line is commented out,
    matrix.go:156-161);
  * LookAtV / Vec3.Normalize / quaternions follow matrix.go:164-177, vector.go,
    quaternion.go.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
FLOAT_CMP_EPSILON = 1e-10  # types/matrix.go:13


def vec3(x, y, z):
    return np.array([x, y, z], dtype=F)


def v_len(v) -> np.float32:
    # float32(math.Sqrt(float64(x*x + y*y + z*z))), sum in float32
    s = F(0)
    acc = None
    for c in v:
        t = F(c) * F(c)
        acc = t if acc is None else F(acc + t)
    del s
    return F(math.sqrt(float(acc)))


def v_normalize(v):
    ln = v_len(v)
    with np.errstate(divide="ignore"):
        inv = F(1.0) / ln
    if inv < FLOAT_CMP_EPSILON:
        return np.zeros(len(v), dtype=F)
    return np.array([F(c) * inv for c in v], dtype=F)


def v_cross(a, b):
    a = [F(x) for x in a]
    b = [F(x) for x in b]
    return np.array(
        [
            F(F(a[1] * b[2]) - F(a[2] * b[1])),
            F(F(a[2] * b[0]) - F(a[0] * b[2])),
            F(F(a[0] * b[1]) - F(a[1] * b[0])),
        ],
        dtype=F,
    )


def v_dot(a, b) -> np.float32:
    acc = None
    for x, y in zip(a, b):
        t = F(F(x) * F(y))
        acc = t if acc is None else F(acc + t)
    return acc


def ident4():
    m = np.zeros(16, dtype=F)
    m[0] = m[5] = m[10] = m[15] = 1
    return m


def translate4(t):
    m = ident4()
    m[12], m[13], m[14] = F(t[0]), F(t[1]), F(t[2])
    return m


def scale4(s):
    s = [F(1.0) if F(x) == 0 else F(x) for x in s]  # matrix.go:42-53
    m = ident4()
    m[0], m[5], m[10] = s
    return m


def mul4(m1, m2):
    out = np.zeros(16, dtype=F)
    for c in range(4):
        for r in range(4):
            acc = None
            for k in range(4):
                t = F(F(m1[r + 4 * k]) * F(m2[4 * c + k]))
                acc = t if acc is None else F(acc + t)
            out[4 * c + r] = acc
    return out


def mul4x1(m, v):
    out = np.zeros(4, dtype=F)
    for r in range(4):
        acc = None
        for k in range(4):
            t = F(F(m[r + 4 * k]) * F(v[k]))
            acc = t if acc is None else F(acc + t)
        out[r] = acc
    return out


# Cofactor tables of Mat4.Inv (matrix.go:108-138).  Each term is sign + element indices,
# products and the running sum are taken left to right exactly as the Go expression.
_DET_TERMS = (
    "+0.5.10.15 -0.5.11.14 -0.6.9.15 +0.6.11.13 +0.7.9.14 -0.7.10.13 "
    "-1.4.10.15 +1.4.11.14 +1.6.8.15 -1.6.11.12 -1.7.8.14 +1.7.10.12 "
    "+2.4.9.15 -2.4.11.13 -2.5.8.15 +2.5.11.12 +2.7.8.13 -2.7.9.12 "
    "-3.4.9.14 +3.4.10.13 +3.5.8.14 -3.5.10.12 -3.6.8.13 +3.6.9.12"
)
_ADJ_TERMS = (
    "-7.10.13 +6.11.13 +7.9.14 -5.11.14 -6.9.15 +5.10.15",
    "+3.10.13 -2.11.13 -3.9.14 +1.11.14 +2.9.15 -1.10.15",
    "-3.6.13 +2.7.13 +3.5.14 -1.7.14 -2.5.15 +1.6.15",
    "+3.6.9 -2.7.9 -3.5.10 +1.7.10 +2.5.11 -1.6.11",
    "+7.10.12 -6.11.12 -7.8.14 +4.11.14 +6.8.15 -4.10.15",
    "-3.10.12 +2.11.12 +3.8.14 -0.11.14 -2.8.15 +0.10.15",
    "+3.6.12 -2.7.12 -3.4.14 +0.7.14 +2.4.15 -0.6.15",
    "-3.6.8 +2.7.8 +3.4.10 -0.7.10 -2.4.11 +0.6.11",
    "-7.9.12 +5.11.12 +7.8.13 -4.11.13 -5.8.15 +4.9.15",
    "+3.9.12 -1.11.12 -3.8.13 +0.11.13 +1.8.15 -0.9.15",
    "-3.5.12 +1.7.12 +3.4.13 -0.7.13 -1.4.15 +0.5.15",
    "+3.5.8 -1.7.8 -3.4.9 +0.7.9 +1.4.11 -0.5.11",
    "+6.9.12 -5.10.12 -6.8.13 +4.10.13 +5.8.14 -4.9.14",
    "-2.9.12 +1.10.12 +2.8.13 -0.10.13 -1.8.14 +0.9.14",
    "+2.5.12 -1.6.12 -2.4.13 +0.6.13 +1.4.14 -0.5.14",
    "-2.5.8 +1.6.8 +2.4.9 -0.6.9 -1.4.10 +0.5.10",
)


def _eval_terms(m, spec: str) -> np.float32:
    acc = None
    for i, term in enumerate(spec.split()):
        sign, idx = term[0], [int(t) for t in term[1:].split(".")]
        if i == 0 and sign == "-":
            # a leading "-m[a]*m[b]*m[c]" negates the first factor, which is exact
            p = F(-F(m[idx[0]]))
        else:
            p = F(m[idx[0]])
        for j in idx[1:]:
            p = F(p * F(m[j]))
        if acc is None:
            acc = p
        elif sign == "+":
            acc = F(acc + p)
        else:
            acc = F(acc - p)
    return acc


def inv4(m):
    det = _eval_terms(m, _DET_TERMS)
    if abs(float(det)) < FLOAT_CMP_EPSILON:
        return np.zeros(16, dtype=F)
    ret = np.array([_eval_terms(m, t) for t in _ADJ_TERMS], dtype=F)
    s = F(F(1) / det)
    return np.array([F(x * s) for x in ret], dtype=F)


def perspective4(fovy, aspect, near, far):
    fovy, aspect, near, far = F(fovy), F(aspect), F(near), F(far)
    nmf = F(near - far)
    f = F(1.0 / math.tan(float(fovy) / 2.0))
    m = np.zeros(16, dtype=F)
    m[0] = F(f / aspect)
    m[5] = f
    m[10] = F(F(near + far) / nmf)
    m[11] = F(-1)
    m[14] = F(F(F(F(2.0) * far) * near) / nmf)
    return m


def look_at_v(eye, center, up):
    eye = np.asarray(eye, dtype=F)
    center = np.asarray(center, dtype=F)
    f = v_normalize(center - eye)
    s = v_normalize(v_cross(f, v_normalize(np.asarray(up, dtype=F))))
    u = v_cross(s, f)
    rot = np.array(
        [s[0], u[0], -f[0], 0, s[1], u[1], -f[1], 0, s[2], u[2], -f[2], 0, 0, 0, 0, 1],
        dtype=F,
    )
    trans = ident4()
    trans[12], trans[13], trans[14] = -eye[0], -eye[1], -eye[2]
    return mul4(rot, trans)


# --- quaternions (types/quaternion.go), only what Camera.Update and the instance parser use
def quat_from_axis_angle(axis, angle):
    half = F(F(angle) * F(0.5))
    s = F(math.sin(float(half)))
    c = F(math.cos(float(half)))
    return (np.array([F(a) * s for a in axis], dtype=F), c)


def quat_mul(q1, q2):
    v1, w1 = q1
    v2, w2 = q2
    v = v_cross(v1, v2) + np.array([F(x) * w1 for x in v2], dtype=F)
    v = (v + np.array([F(x) * w2 for x in v1], dtype=F)).astype(F)
    w = F(F(w1 * w2) - v_dot(v1, v2))
    return (v, w)


def quat_normalize(q):
    v, w = q
    acc = F(w * w)
    for c in v:
        acc = F(acc + F(c * c))
    length = F(math.sqrt(float(acc)))
    if abs(1.0 - float(length)) < FLOAT_CMP_EPSILON:
        return q
    if length == 0:
        return (np.zeros(3, dtype=F), F(1))
    inv = F(F(1) / length)
    return (np.array([F(c * inv) for c in v], dtype=F), F(F(w * F(1)) / length))


def quat_rotate(q, vec):
    v, w = q
    cross = v_cross(v, vec)
    two_w = F(F(2) * w)
    a = (np.asarray(vec, dtype=F) + np.array([F(c * two_w) for c in cross], dtype=F)).astype(F)
    return (a + v_cross(np.array([F(c * F(2)) for c in v], dtype=F), cross)).astype(F)


def quat_mat4(q):
    v, w = q
    x, y, z = v
    two = F(2)

    def t(a, b):
        return F(F(two * a) * b)

    one = F(1)
    return np.array(
        [
            F(F(one - t(y, y)) - t(z, z)), F(t(x, y) + t(w, z)), F(t(x, z) - t(w, y)), 0,
            F(t(x, y) - t(w, z)), F(F(one - t(x, x)) - t(z, z)), F(t(y, z) + t(w, x)), 0,
            F(t(x, z) + t(w, y)), F(t(y, z) - t(w, x)), F(F(one - t(x, x)) - t(y, y)), 0,
            0, 0, 0, 1,
        ],
        dtype=F,
    )
