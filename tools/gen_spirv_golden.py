#!/usr/bin/env python
"""Generates tests/golden/frag_spirv_vectors.npz and vert_spirv_vectors.npz by EXECUTING the
reference's own compiled shaders (shaders/trace.frag.spv, shaders/trace.vert.spv) with
tools/spirv_interp.py.  Run here, in the authoring container, where /root/reference exists:

    python tools/gen_spirv_golden.py

The vectors are committed; tests/test_oracle_spirv.py checks the oracle against them and never
touches /root/reference.  Every vector records the inputs of one shader invocation and what the
shader binary produced: discard flag, `color`, gl_FragDepth and the values of the local variables
`model_ray_voxel`, `steps`, `mask` at exit (found through the module's OpName debug info).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import scenes  # noqa: E402
from tools import spirv_interp as S  # noqa: E402
from vtrace_b200 import glm  # noqa: E402

REF = "/root/reference/shaders"
F = np.float32


def cols(a):
    a = np.asarray(a, dtype=F).reshape(4, 4)
    return [a[c].copy() for c in range(4)]


def synthetic_volume(seed, w, h, d, fill, alphas=(255,)):
    rng = np.random.default_rng(seed)
    vol = np.zeros((d, h, w, 4), dtype=np.uint8)
    filled = rng.random((d, h, w)) < fill
    vol[..., :3] = rng.integers(0, 256, size=(d, h, w, 3), dtype=np.uint8)
    vol[..., 3] = np.where(filled, rng.choice(np.array(alphas, dtype=np.uint8), size=(d, h, w)), 0)
    vol[~filled] = 0
    return vol.reshape(-1)


def face_point(rng, eye_m):
    """A random point on a face of the unit cube that faces the (model-space) eye."""
    faces = [(k, s) for k in range(3) for s in (-0.5, 0.5) if (eye_m[k] - s) * s > 0]
    k, s = faces[rng.integers(len(faces))]
    p = rng.uniform(-0.5, 0.5, size=3).astype(F)
    if rng.random() < 0.15:  # hug an edge / corner
        j = (k + 1 + rng.integers(2)) % 3
        p[j] = F(rng.choice([-0.5, 0.5]) * (1.0 - 10.0 ** rng.uniform(-7, -2)))
    p[k] = F(s)
    return p


def main():
    frag = S.Module(os.path.join(REF, "trace.frag.spv"))
    vert = S.Module(os.path.join(REF, "trace.vert.spv"))
    sha = {n: hashlib.sha256(open(os.path.join(REF, n), "rb").read()).hexdigest() for n in ("trace.frag.spv", "trace.vert.spv")}

    tex_names = ["Treasure", "AncientTemple", "synthetic_22x23x29", "synthetic_1x1x1", "synthetic_alpha_12"]
    t0, t1 = scenes.load_asset("Treasure"), scenes.load_asset("AncientTemple")
    raws = [t0.get_raw(), t1.get_raw(), synthetic_volume(3, 22, 23, 29, 0.12), np.array([9, 200, 30, 255], dtype=np.uint8),
            synthetic_volume(5, 12, 12, 12, 0.25, alphas=(255, 128, 1))]
    dims = [t0.dims(), t1.dims(), (22, 23, 29), (1, 1, 1), (12, 12, 12)]
    textures = [S.Texture3D(r, *d) for r, d in zip(raws, dims)]

    rng = np.random.default_rng(20261017)
    rows = []
    models = [glm.identity(),
              glm.translate(glm.identity(), (0.3, -0.2, 0.1)),
              glm.scale(glm.rotate(glm.translate(glm.identity(), (0.4, 0.1, -0.3)), 0.7, (0.3, 1.0, 0.2)), (1.3, 0.7, 1.9)),
              glm.scale(glm.identity(), (2.0, 2.0, 2.0))]
    eyes = [scenes.EYE, (0.8, -0.45, 0.6), (-0.7, 0.2, 0.9), (0.0, -1.2, 0.05), (2.5, 2.0, -1.7), (0.0, 0.0, 2.0), (2.0, 0.0, 0.0)]
    plan = []
    for ti in range(len(textures)):
        for mi, M in enumerate(models):
            for ei, eye in enumerate(eyes):
                n = 6 if ti < 2 else 3
                plan.append((ti, mi, ei, n))
    for ti, mi, ei, n in plan:
        M = np.asarray(models[mi], dtype=F)
        eye = np.asarray(eyes[ei], dtype=F)
        P = glm.perspective(glm.REFERENCE_FOV, F(16.0 / 9.0), glm.REFERENCE_NEAR, glm.REFERENCE_FAR)
        center = np.asarray(M[3][:3], dtype=F)
        V = glm.look_at(eye, center, scenes.UP)
        Mi = np.linalg.inv(M.astype(np.float64).T)  # row-major inverse, only to find visible faces
        eye_m = (Mi @ np.append(eye.astype(np.float64), 1.0))[:3]
        if np.all(np.abs(eye_m) <= 0.5):
            continue
        PV = S.mat_times_mat(cols(P), cols(V))
        for k in range(n):
            if ei >= 5 and k == 0:
                # exactly axis-aligned ray through the centre of the facing face (zero direction components)
                axis = int(np.argmax(np.abs(eye_m)))
                mp3 = np.zeros(3, dtype=F)
                mp3[axis] = F(0.5 if eye_m[axis] > 0 else -0.5)
            else:
                mp3 = face_point(rng, eye_m)
            mp = np.append(mp3, F(1.0)).astype(F)
            world = S.mat_times_vec(cols(M), mp)          # trace.vert:44
            sp = S.mat_times_vec(PV, world)               # trace.vert:45
            inv = S.Invocation(frag, {"screen_position": sp, "world_position": world, "model_position": mp,
                                      "object_id": np.uint32(0), "texture_id": np.uint32(ti), "model_matrix": cols(M),
                                      "push": [cols(P), cols(V)]}, textures=textures)
            inv.run()
            color = np.zeros(4, dtype=F) if inv.discarded else np.asarray(inv.output("color"), dtype=F)
            rows.append(dict(P=P.reshape(16), V=V.reshape(16), M=M.reshape(16), sp=sp, mp=mp, tex=ti,
                             discard=int(inv.discarded), color=color, depth=F(inv.output("gl_FragDepth")),
                             voxel=np.asarray(inv.local("model_ray_voxel"), dtype=np.int32),
                             steps=int(inv.local("steps")),
                             mask=np.asarray(inv.locals_by_name.get("mask", [np.zeros(3, bool)])[0], dtype=np.uint8)))
    out = {k: np.stack([np.asarray(r[k]) for r in rows]) for k in rows[0]}
    out["tex_names"] = np.array(tex_names)
    out["tex_dims"] = np.array(dims, dtype=np.int32)
    for i in (2, 3, 4):
        out[f"tex_raw_{i}"] = raws[i]
    out["spv_sha256"] = np.array([sha["trace.frag.spv"]])
    path = os.path.join(ROOT, "tests", "golden", "frag_spirv_vectors.npz")
    np.savez_compressed(path, **out)
    n_hit = int((out["discard"] == 0).sum())
    print(f"{path}: {len(rows)} fragments, {n_hit} hits, {len(rows) - n_hit} discards, max steps {out['steps'].max()}")

    # ---- vertex shader: 8 cube corners x a few instances ------------------------------------
    vrows = []
    corners = [(-0.5, -0.5, 0.5), (0.5, -0.5, 0.5), (-0.5, 0.5, 0.5), (0.5, 0.5, 0.5),
               (-0.5, -0.5, -0.5), (0.5, -0.5, -0.5), (-0.5, 0.5, -0.5), (0.5, 0.5, -0.5)]  # lib/memory.c:22-31
    P, V = scenes.camera(1920, 1080)
    for mi, M in enumerate(models):
        for tid in (0, 1, 77, 65535):
            inst = glm.with_texture_id(M, tid)
            for c in corners:
                inv = S.Invocation(vert, {"position": np.array(c, dtype=F), "model": cols(inst), "push": [cols(P), cols(V)],
                                          "gl_InstanceIndex": np.int32(mi)})
                inv.run()
                vrows.append(dict(P=P.reshape(16), V=V.reshape(16), inst=np.asarray(inst, dtype=F).reshape(16),
                                  position=np.array(c, dtype=F), screen_position=np.asarray(inv.output("screen_position"), dtype=F),
                                  model_position=np.asarray(inv.output("model_position"), dtype=F),
                                  texture_id=np.uint32(inv.output("texture_id")),
                                  model_matrix=np.concatenate(inv.output("model_matrix")).astype(F)))
    vout = {k: np.stack([np.asarray(r[k]) for r in vrows]) for k in vrows[0]}
    vout["spv_sha256"] = np.array([sha["trace.vert.spv"]])
    vpath = os.path.join(ROOT, "tests", "golden", "vert_spirv_vectors.npz")
    np.savez_compressed(vpath, **vout)
    print(f"{vpath}: {len(vrows)} vertices")


if __name__ == "__main__":
    main()
