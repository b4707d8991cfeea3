#!/usr/bin/env python
"""A small SPIR-V interpreter, just large enough to EXECUTE the reference's own compiled shaders
(shaders/trace.frag.spv, shaders/trace.vert.spv) one invocation at a time on the CPU.

Why: the reference cannot run in this image (no Vulkan ICD, GLFW, rustc) and ships no golden
vectors for the traversal path.  Running its committed shader binary through
this interpreter is the closest thing to "outputs of the reference itself": it pins the oracle's
transcription of trace.frag (operation order, tie rule `side <= min(other two)`, lower-edge texel
coordinate, loop bound, discard) against the bytes the reference's build produced.

All floating point is IEEE binary32 with one rounding per SPIR-V instruction.  Where SPIR-V leaves
precision or NaN behaviour to the implementation the choices are spelled out below and shared with
the oracle: OpMatrixTimesVector / OpMatrixTimesMatrix accumulate over columns left to right;
MatrixInverse is the cofactor expansion of oracle/vtrace_oracle.c; Normalize = x / Length;
FMin is minNum; the NEAREST clamp-to-edge sampler picks texel floor(u * size).

Only tools/gen_spirv_golden.py imports this (at generation time, reading /root/reference); tests
read the generated vectors, never the reference tree.
"""
from __future__ import annotations

import struct

import numpy as np

F = np.float32


# ---------------------------------------------------------------------------------------------
# value helpers

def f32(x):
    return np.float32(x)


def mat_inverse(cols):
    """cols: 4 column vectors (np.float32[4]).  Same cofactor expansion / order as the oracle."""
    with np.errstate(all="ignore"):
        return _mat_inverse(cols)


def _mat_inverse(cols):
    A = lambda r, c: cols[c][r]  # noqa: E731
    s0 = A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1)
    s1 = A(0, 0) * A(1, 2) - A(1, 0) * A(0, 2)
    s2 = A(0, 0) * A(1, 3) - A(1, 0) * A(0, 3)
    s3 = A(0, 1) * A(1, 2) - A(1, 1) * A(0, 2)
    s4 = A(0, 1) * A(1, 3) - A(1, 1) * A(0, 3)
    s5 = A(0, 2) * A(1, 3) - A(1, 2) * A(0, 3)
    c5 = A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3)
    c4 = A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3)
    c3 = A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2)
    c2 = A(2, 0) * A(3, 3) - A(3, 0) * A(2, 3)
    c1 = A(2, 0) * A(3, 2) - A(3, 0) * A(2, 2)
    c0 = A(2, 0) * A(3, 1) - A(3, 0) * A(2, 1)
    det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0
    idet = F(1.0) / det
    B = np.zeros((4, 4), dtype=F)  # B[r][c]
    B[0][0] = ((A(1, 1) * c5 - A(1, 2) * c4) + A(1, 3) * c3) * idet
    B[0][1] = ((-A(0, 1) * c5 + A(0, 2) * c4) - A(0, 3) * c3) * idet
    B[0][2] = ((A(3, 1) * s5 - A(3, 2) * s4) + A(3, 3) * s3) * idet
    B[0][3] = ((-A(2, 1) * s5 + A(2, 2) * s4) - A(2, 3) * s3) * idet
    B[1][0] = ((-A(1, 0) * c5 + A(1, 2) * c2) - A(1, 3) * c1) * idet
    B[1][1] = ((A(0, 0) * c5 - A(0, 2) * c2) + A(0, 3) * c1) * idet
    B[1][2] = ((-A(3, 0) * s5 + A(3, 2) * s2) - A(3, 3) * s1) * idet
    B[1][3] = ((A(2, 0) * s5 - A(2, 2) * s2) + A(2, 3) * s1) * idet
    B[2][0] = ((A(1, 0) * c4 - A(1, 1) * c2) + A(1, 3) * c0) * idet
    B[2][1] = ((-A(0, 0) * c4 + A(0, 1) * c2) - A(0, 3) * c0) * idet
    B[2][2] = ((A(3, 0) * s4 - A(3, 1) * s2) + A(3, 3) * s0) * idet
    B[2][3] = ((-A(2, 0) * s4 + A(2, 1) * s2) - A(2, 3) * s0) * idet
    B[3][0] = ((-A(1, 0) * c3 + A(1, 1) * c1) - A(1, 2) * c0) * idet
    B[3][1] = ((A(0, 0) * c3 - A(0, 1) * c1) + A(0, 2) * c0) * idet
    B[3][2] = ((-A(3, 0) * s3 + A(3, 1) * s1) - A(3, 2) * s0) * idet
    B[3][3] = ((A(2, 0) * s3 - A(2, 1) * s1) + A(2, 2) * s0) * idet
    return [np.array([B[r][c] for r in range(4)], dtype=F) for c in range(4)]


def mat_times_vec(cols, v):
    n = len(cols[0])
    out = np.zeros(n, dtype=F)
    with np.errstate(all="ignore"):
        for i in range(n):
            acc = cols[0][i] * v[0]
            for c in range(1, len(cols)):
                acc = F(acc + F(cols[c][i] * v[c]))
            out[i] = acc
    return out


def mat_times_mat(a, b):
    return [mat_times_vec(a, b[j]) for j in range(len(b))]


def fmin(a, b):
    """GLSL.std.450 FMin with minNum NaN behaviour (the oracle's vo_fmin)."""
    a, b = np.asarray(a, dtype=F), np.asarray(b, dtype=F)
    return np.where(np.isnan(a), b, np.where(np.isnan(b), a, np.where(a < b, a, b))).astype(F)


def srgb_decode_table():
    def lin(c):
        return c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4
    return np.array([lin(k / 255.0) for k in range(256)], dtype=np.float64).astype(F)


class Texture3D:
    """VK_FORMAT_R8G8B8A8_SRGB 3D image + NEAREST / CLAMP_TO_EDGE sampler (lib/descriptor.c:97-117)."""

    _dec = None

    def __init__(self, rgba: np.ndarray, w: int, h: int, d: int):
        self.w, self.h, self.d = w, h, d
        self.texels = np.asarray(rgba, dtype=np.uint8).reshape(d, h, w, 4)  # x fastest (lib/memory.c:353-366)
        if Texture3D._dec is None:
            Texture3D._dec = srgb_decode_table()

    def size(self):
        return np.array([self.w, self.h, self.d], dtype=np.int32)

    def sample(self, uvw):
        idx = []
        with np.errstate(all="ignore"):
            for u, n in zip(uvw, (self.w, self.h, self.d)):
                t = np.floor(F(F(u) * F(n)))
                i = 0 if np.isnan(t) else int(min(max(t, -2**31), 2**31 - 1))
                idx.append(min(max(i, 0), n - 1))
        r, g, b, a = self.texels[idx[2], idx[1], idx[0]]
        dec = Texture3D._dec
        return np.array([dec[r], dec[g], dec[b], F(a) / F(255.0)], dtype=F)


# ---------------------------------------------------------------------------------------------
# module parsing

class Module:
    def __init__(self, path: str):
        data = open(path, "rb").read()
        self.words = struct.unpack("<%dI" % (len(data) // 4), data)
        assert self.words[0] == 0x07230203, "not SPIR-V"
        self.version = self.words[1]
        self.names, self.member_names = {}, {}
        self.types, self.consts, self.decor = {}, {}, {}
        self.global_vars = {}   # id -> (type id, storage class)
        self.ext_sets = {}
        self.functions = {}     # id -> list of (op, operands)
        self.entry = None
        self._parse()

    @staticmethod
    def _string(ws):
        b = b"".join(struct.pack("<I", w) for w in ws)
        return b.split(b"\0")[0].decode()

    def _parse(self):
        w, i, cur = self.words, 5, None
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            o = w[i + 1:i + wc]
            i += wc
            if op == 5:
                self.names[o[0]] = self._string(o[1:])
            elif op == 6:
                self.member_names[(o[0], o[1])] = self._string(o[2:])
            elif op == 11:
                self.ext_sets[o[0]] = self._string(o[1:])
            elif op == 15:
                self.entry = o[1]
            elif op == 71:
                self.decor.setdefault(o[0], []).append(o[1:])
            elif op == 19:
                self.types[o[0]] = ("void",)
            elif op == 20:
                self.types[o[0]] = ("bool",)
            elif op == 21:
                self.types[o[0]] = ("int", o[1], o[2])
            elif op == 22:
                self.types[o[0]] = ("float", o[1])
            elif op == 23:
                self.types[o[0]] = ("vector", o[1], o[2])
            elif op == 24:
                self.types[o[0]] = ("matrix", o[1], o[2])
            elif op == 25:
                self.types[o[0]] = ("image",)
            elif op == 27:
                self.types[o[0]] = ("sampled_image", o[1])
            elif op == 28:
                self.types[o[0]] = ("array", o[1], o[2])
            elif op == 29:
                self.types[o[0]] = ("runtime_array", o[1])
            elif op == 30:
                self.types[o[0]] = ("struct", list(o[1:]))
            elif op == 32:
                self.types[o[0]] = ("pointer", o[1], o[2])
            elif op == 33:
                self.types[o[0]] = ("function",)
            elif op == 43:
                self.consts[o[1]] = self._scalar_const(o[0], o[2])
            elif op == 44:
                self.consts[o[1]] = self._composite_const(o[0], [self.consts[x] for x in o[2:]])
            elif op == 59 and cur is None:
                self.global_vars[o[1]] = (o[0], o[2])
            elif op == 54:
                cur = o[1]
                self.functions[cur] = []
            elif op == 56:
                cur = None
            elif cur is not None:
                self.functions[cur].append((op, o))

    def _scalar_const(self, ty, word):
        t = self.types[ty]
        if t[0] == "float":
            return np.frombuffer(struct.pack("<I", word), dtype=F)[0]
        if t[0] == "int":
            return np.int32(struct.unpack("<i", struct.pack("<I", word))[0]) if t[2] else np.uint32(word)
        raise NotImplementedError(t)

    def _composite_const(self, ty, parts):
        t = self.types[ty]
        if t[0] == "vector":
            return np.array(parts)
        if t[0] == "matrix":
            return [np.array(p, dtype=F) for p in parts]
        raise NotImplementedError(t)

    def zero_value(self, ty):
        t = self.types[ty]
        if t[0] == "float":
            return F(0)
        if t[0] == "int":
            return np.int32(0) if t[2] else np.uint32(0)
        if t[0] == "bool":
            return np.bool_(False)
        if t[0] == "vector":
            return np.array([self.zero_value(t[1]) for _ in range(t[2])])
        if t[0] == "matrix":
            return [self.zero_value(t[1]) for _ in range(t[2])]
        if t[0] == "struct":
            return [self.zero_value(m) for m in t[1]]
        if t[0] == "array":
            return [self.zero_value(t[1]) for _ in range(int(self.consts[t[2]]))]
        raise NotImplementedError(t)


# ---------------------------------------------------------------------------------------------
# execution

class Pointer:
    def __init__(self, cell, path=()):
        self.cell, self.path = cell, tuple(path)  # cell: single-element list holding the object

    def load(self):
        v = self.cell[0]
        for p in self.path:
            v = v[p]
        return v

    def store(self, val):
        if not self.path:
            self.cell[0] = val
            return
        v = self.cell[0]
        for p in self.path[:-1]:
            v = v[p]
        v[self.path[-1]] = val


class Discard(Exception):
    pass


def _copy(v):
    if isinstance(v, np.ndarray):
        return v.copy()
    if isinstance(v, list):
        return [_copy(x) for x in v]
    return v


class Invocation:
    """One shader invocation.  `inputs`/`outputs` are addressed by OpName."""

    def __init__(self, module: Module, inputs: dict, textures: list | None = None, max_instructions: int = 2_000_000):
        self.m = module
        self.textures = textures or []
        self.val = dict(module.consts)
        self.cells = {}
        self.max_instructions = max_instructions
        for vid, (pty, storage) in module.global_vars.items():
            name = module.names.get(vid, "")
            pointee = module.types[pty][2]
            if module.types[pointee][0] == "runtime_array":  # sampler3D tex[]
                cell = [self.textures]
            elif name in inputs:
                cell = [_copy(inputs[name])]
            elif name == "" and storage == 9 and "push" in inputs:  # unnamed push-constant block instance
                cell = [_copy(inputs["push"])]
            else:
                cell = [module.zero_value(pointee)]
            self.cells[vid] = cell
            self.val[vid] = Pointer(cell)
        self.locals_by_name = {}
        self.discarded = False
        self.executed = 0

    def output(self, name):
        for vid, _ in self.m.global_vars.items():
            if self.m.names.get(vid) == name:
                return self.cells[vid][0]
        raise KeyError(name)

    def local(self, name):
        return self.locals_by_name[name][0]

    # -- helpers -------------------------------------------------------------------------------
    def _ext(self, inst, args):
        a = [self.val[x] for x in args]
        with np.errstate(all="ignore"):
            if inst == 34:   # MatrixInverse
                return mat_inverse(a[0])
            if inst == 69:   # Normalize
                ln = self._length(a[0])
                return (a[0] / ln).astype(F)
            if inst == 66:   # Length
                return self._length(a[0])
            if inst == 6:    # FSign
                x = np.asarray(a[0], dtype=F)
                return np.where(x > 0, F(1), np.where(x < 0, F(-1), F(0))).astype(F)
            if inst == 37:   # FMin
                return fmin(a[0], a[1])
            if inst == 39:   # SMin
                return np.minimum(a[0], a[1]).astype(np.int32)
            if inst == 8:    # Floor
                return np.floor(a[0]).astype(F)
            if inst == 4:    # FAbs
                return np.abs(a[0]).astype(F)
        raise NotImplementedError(f"GLSL.std.450 instruction {inst}")

    @staticmethod
    def _length(v):
        v = np.asarray(v, dtype=F)
        with np.errstate(all="ignore"):
            acc = F(v[0] * v[0])
            for k in range(1, len(v)):
                acc = F(acc + F(v[k] * v[k]))
            return F(np.sqrt(acc))

    @staticmethod
    def _f2s(x):
        x = np.asarray(x, dtype=F)
        with np.errstate(all="ignore"):
            t = np.where(np.isnan(x), F(0), np.trunc(x))
            t = np.clip(t, -2147483648.0, 2147483647.0)
        return t.astype(np.int64).astype(np.int32)

    # -- main loop -----------------------------------------------------------------------------
    def run(self, fn_id=None):
        m = self.m
        body = m.functions[fn_id or m.entry]
        labels = {o[0]: k for k, (op, o) in enumerate(body) if op == 248}
        pc, cur_label, prev_label = 0, None, None
        prev_label_for_phi = None
        V = self.val
        with np.errstate(all="ignore"):
            while pc < len(body):
                op, o = body[pc]
                pc += 1
                self.executed += 1
                if self.executed > self.max_instructions:
                    raise RuntimeError("instruction budget exceeded")
                if op == 248:      # Label
                    prev_label, cur_label = cur_label, o[0]
                elif op == 59:     # Variable (Function storage)
                    pointee = m.types[o[0]][2]
                    cell = [_copy(V[o[3]]) if len(o) > 3 else m.zero_value(pointee)]
                    V[o[1]] = Pointer(cell)
                    self.locals_by_name[m.names.get(o[1], str(o[1]))] = cell
                elif op == 61:     # Load
                    V[o[1]] = _copy(V[o[2]].load())
                elif op == 62:     # Store
                    V[o[0]].store(_copy(V[o[1]]))
                elif op == 65:     # AccessChain
                    base = V[o[2]]
                    idx = tuple(int(V[x]) for x in o[3:])
                    V[o[1]] = Pointer(base.cell, base.path + idx)
                elif op == 79:     # VectorShuffle
                    cat = np.concatenate([np.atleast_1d(V[o[2]]), np.atleast_1d(V[o[3]])])
                    V[o[1]] = np.array([cat[k] for k in o[4:]])
                elif op == 80:     # CompositeConstruct
                    t = m.types[o[0]]
                    parts = [V[x] for x in o[2:]]
                    if t[0] == "vector":
                        V[o[1]] = np.concatenate([np.atleast_1d(p) for p in parts])
                    elif t[0] == "matrix":
                        V[o[1]] = [np.array(p, dtype=F) for p in parts]
                    else:
                        raise NotImplementedError(t)
                elif op == 81:     # CompositeExtract
                    v = V[o[2]]
                    for k in o[3:]:
                        v = v[k]
                    V[o[1]] = _copy(v)
                elif op == 12:     # ExtInst
                    V[o[1]] = self._ext(o[3], o[4:])
                elif op == 87:     # ImageSampleImplicitLod
                    V[o[1]] = V[o[2]].sample(V[o[3]])
                elif op == 100:    # Image
                    V[o[1]] = V[o[2]]
                elif op == 103:    # ImageQuerySizeLod
                    V[o[1]] = V[o[2]].size()
                elif op == 110:    # ConvertFToS
                    V[o[1]] = self._f2s(V[o[2]])
                elif op == 111:    # ConvertSToF
                    V[o[1]] = np.asarray(V[o[2]]).astype(F)
                elif op == 124:    # Bitcast
                    src = np.asarray(V[o[2]])
                    t = m.types[o[0]]
                    base = m.types[t[1]] if t[0] == "vector" else t
                    dt = F if base[0] == "float" else (np.int32 if base[2] else np.uint32)
                    V[o[1]] = src.view(dt) if src.ndim else np.array([src]).view(dt)[0]
                elif op == 128:    # IAdd
                    a, b = np.asarray(V[o[2]]), np.asarray(V[o[3]])
                    V[o[1]] = (a.astype(np.int64) + b.astype(np.int64)).astype(a.dtype)
                elif op == 132:    # IMul
                    a, b = np.asarray(V[o[2]]), np.asarray(V[o[3]])
                    V[o[1]] = (a.astype(np.int64) * b.astype(np.int64)).astype(a.dtype)
                elif op == 129:
                    V[o[1]] = (np.asarray(V[o[2]], dtype=F) + np.asarray(V[o[3]], dtype=F)).astype(F)
                elif op == 131:
                    V[o[1]] = (np.asarray(V[o[2]], dtype=F) - np.asarray(V[o[3]], dtype=F)).astype(F)
                elif op == 133:
                    V[o[1]] = (np.asarray(V[o[2]], dtype=F) * np.asarray(V[o[3]], dtype=F)).astype(F)
                elif op == 136:
                    V[o[1]] = (np.asarray(V[o[2]], dtype=F) / np.asarray(V[o[3]], dtype=F)).astype(F)
                elif op == 142:    # VectorTimesScalar
                    V[o[1]] = (np.asarray(V[o[2]], dtype=F) * F(V[o[3]])).astype(F)
                elif op == 145:    # MatrixTimesVector
                    V[o[1]] = mat_times_vec(V[o[2]], V[o[3]])
                elif op == 146:    # MatrixTimesMatrix
                    V[o[1]] = mat_times_mat(V[o[2]], V[o[3]])
                elif op == 154:    # Any
                    V[o[1]] = np.bool_(np.any(V[o[2]]))
                elif op == 155:    # All
                    V[o[1]] = np.bool_(np.all(V[o[2]]))
                elif op == 167:    # LogicalAnd
                    V[o[1]] = np.logical_and(V[o[2]], V[o[3]])
                elif op == 169:    # Select
                    V[o[1]] = np.where(V[o[2]], V[o[3]], V[o[4]])
                    if np.ndim(V[o[1]]) == 0:
                        V[o[1]] = V[o[1]][()]
                elif op in (175, 177, 179, 173, 176, 172, 174, 178, 170, 171):  # integer compares
                    a, b = np.asarray(V[o[2]]), np.asarray(V[o[3]])
                    if op in (176, 172, 174, 178):
                        a, b = a.astype(np.uint32), b.astype(np.uint32)
                    fn = {175: np.greater_equal, 177: np.less, 179: np.less_equal, 173: np.greater, 176: np.less,
                          172: np.greater, 174: np.greater_equal, 178: np.less_equal, 170: np.equal, 171: np.not_equal}[op]
                    V[o[1]] = fn(a, b)
                elif op in (180, 182, 184, 186, 188, 190):  # ordered float compares
                    a, b = np.asarray(V[o[2]], dtype=F), np.asarray(V[o[3]], dtype=F)
                    fn = {180: np.equal, 182: np.not_equal, 184: np.less, 186: np.greater, 188: np.less_equal,
                          190: np.greater_equal}[op]
                    res = fn(a, b)
                    if op == 182:
                        res = np.logical_and(res, ~(np.isnan(a) | np.isnan(b)))
                    V[o[1]] = res
                elif op == 245:    # Phi
                    for k in range(2, len(o), 2):
                        if o[k + 1] == prev_label_for_phi:
                            V[o[1]] = _copy(V[o[k]])
                            break
                    else:
                        raise RuntimeError("phi: no incoming edge matches")
                elif op in (246, 247):  # LoopMerge / SelectionMerge
                    pass
                elif op == 249:    # Branch
                    prev_label_for_phi = cur_label
                    pc = labels[o[0]]
                elif op == 250:    # BranchConditional
                    prev_label_for_phi = cur_label
                    pc = labels[o[1]] if bool(V[o[0]]) else labels[o[2]]
                elif op == 252:    # Kill
                    self.discarded = True
                    return self
                elif op == 253:    # Return
                    return self
                else:
                    raise NotImplementedError(f"SPIR-V opcode {op}")
        return self
