"""CPU restatement of the camera observations (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Follows what the reference asks MuJoCo for at gym_kmanip/env_sim.py:140-145 (``physics.render(height, width, camera_id)``
per camera of ``obs_list``) and env_sim.py:187-188 (``k_render``), restricted to what can exist in this repository: the
reference renders with MuJoCo's OpenGL pipeline and its visual meshes are absent from the snapshot, so -- like the CUDA
path (gym_kmanip_b200/csrc/km_render.cuh) -- this draws the completed model's primitives by ray casting under MuJoCo's
camera conventions (mj_camlight ``targetbody``: z = normalize(cam - target), x = normalize(up x z), y = z x x; looks along
-z; vertical ``fovy``) and fixed-function lighting model (Blinn-Phong, headlight at the camera + the directional lights of
scene.xml:10-12, material = geom rgba, specular 0.5, exponent 0.5 * 128).  PARITY UNPINNED: there is no MuJoCo image to
compare with; the tests compare the CUDA path with this restatement.

Stage 1 (``scene_record``) works from MuJoCo-style *body* frames (the C++ oracle's xpos / xquat, fp64); stage 2
(``render_record``) is plain numpy float32 over all pixels and all primitives (no culling).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

REC_HDR, PRIM_FLOATS = 16, 16
SPHERE, CAPSULE, BOX = 0, 1, 2
MAT_TABLE, MAT_CUBE, MAT_LINK, MAT_PAD = 0, 1, 2, 3
JNT_FREE = 0
GEOM_PLANE, GEOM_SPHERE, GEOM_BOX = 0, 2, 6


def _q2mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def _unit(v):
    return v / np.linalg.norm(v)


def params(flat: Dict, cam_name: str, width: int, height: int) -> Dict:
    """Camera intrinsics, lights and colours of one camera, as float32-rounded numbers."""
    ci = flat["cam_name"].index(cam_name)
    f32 = lambda a: np.asarray(a, dtype=np.float32)   # noqa: E731
    gi_table = flat["geom_type"].index(GEOM_PLANE)
    tb = flat["geom_bodyid"][gi_table]
    # the table body is static: world z of the plane = chain of body offsets (identity orientations in the scenes)
    z, b = flat["geom_pos"][gi_table][2], tb
    while b != 0:
        z += flat["body_pos"][b][2]
        b = flat["body_parent"][b]
    vis = flat["visual"]
    k = 0
    while (1 << (k + 1)) <= int(vis["shininess"] * 128.0 + 0.5):
        k += 1
    amb = np.array(flat["headlight"]["ambient"], dtype=np.float64) + np.sum(np.array(flat["light_ambient"]).reshape(-1, 3), axis=0)
    return dict(
        W=width, H=height, focal=np.float32(0.5 * height / math.tan(0.5 * flat["cam_fovy"][ci] * math.pi / 180.0)),
        tab_z=np.float32(z), ambient=f32(amb), head_diffuse=f32(flat["headlight"]["diffuse"]),
        head_specular=f32(flat["headlight"]["specular"]),
        ldir=f32([-_unit(np.array(d)) for d in flat["light_dir"]]), ldiffuse=f32(flat["light_diffuse"]),
        lspecular=f32(flat["light_specular"]),
        mat=f32([flat["geom_rgba"][gi_table][:3], flat["geom_rgba"][flat["geom_name"].index("cube")][:3], vis["link_rgba"][:3],
                 flat["geom_rgba"][flat["geom_type"].index(GEOM_SPHERE)][:3]]),
        mat_specular=np.float32(vis["specular"]), squarings=k, spec_cut=np.float32(2.0 ** (-20.0 / (1 << k))), link_radius=np.float32(vis["link_radius"]), cam=ci)


def scene_record(flat: Dict, xpos: np.ndarray, xquat: np.ndarray, cam_name: str) -> np.ndarray:
    """Render record [camera origin, x, y, z axes, nprim | primitives] from body frames (xpos [nbody, 3], xquat [nbody, 4])."""
    ci = flat["cam_name"].index(cam_name)
    xpos, xquat = np.asarray(xpos, float).reshape(-1, 3), np.asarray(xquat, float).reshape(-1, 4)
    cb, tb = flat["cam_bodyid"][ci], flat["cam_targetbodyid"][ci]
    o = xpos[cb] + _q2mat(xquat[cb]) @ np.array(flat["cam_pos"][ci])
    z = _unit(o - xpos[tb])
    x = _unit(np.cross([0.0, 0.0, 1.0], z))
    y = _unit(np.cross(z, x))
    prims = []
    gi = flat["geom_name"].index("cube")
    b = flat["geom_bodyid"][gi]
    prims.append([BOX | (MAT_CUBE << 4)] + list(xpos[b]) + list(flat["geom_size"][gi]) + list(_q2mat(xquat[b]).reshape(-1)))
    for gi in range(flat["ngeom"]):          # finger pads in geom order (= pad order of the device model)
        if flat["geom_type"][gi] != GEOM_SPHERE:
            continue
        b = flat["geom_bodyid"][gi]
        c = xpos[b] + _q2mat(xquat[b]) @ np.array(flat["geom_pos"][gi])
        prims.append([SPHERE | (MAT_PAD << 4)] + list(c) + [flat["geom_size"][gi][0]] + [0.0] * 11)
    r = flat["visual"]["link_radius"]
    cl = cb                                   # the moving body the camera rides on (0: none)
    while cl != 0 and flat["body_jntnum"][cl] == 0:
        cl = flat["body_parent"][cl]
    for j in range(flat["njnt"]):            # one capsule per moving link with a moving parent link, in joint order
        if flat["jnt_type"][j] == JNT_FREE:
            continue
        b = flat["jnt_bodyid"][j]
        p = flat["body_parent"][b]
        while p != 0 and flat["body_jntnum"][p] == 0:
            p = flat["body_parent"][p]
        if p == 0:
            continue
        a_, b_ = xpos[p], xpos[b]
        kind = SPHERE if float(np.sum((a_ - b_) ** 2)) < 1e-10 else CAPSULE
        # a camera does not see the proxies that end at its own link (they would wall it in): radius 0 is never hit
        hidden = cl != 0 and (b == cl or p == cl)
        prims.append([(SPHERE if hidden else kind) | (MAT_LINK << 4)] + list(a_) + [0.0 if hidden else r] + list(b_) + [0.0] * 8)
    rec = np.zeros(REC_HDR + PRIM_FLOATS * len(prims), dtype=np.float32)
    rec[0:3], rec[3:6], rec[6:9], rec[9:12] = o, x, y, z
    rec[12:13].view(np.int32)[0] = len(prims)
    for i, p in enumerate(prims):
        blk = rec[REC_HDR + PRIM_FLOATS * i: REC_HDR + PRIM_FLOATS * (i + 1)]
        blk[1:] = np.asarray(p[1:], dtype=np.float32)
        blk[0:1].view(np.int32)[0] = p[0]
    return rec


def _dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def _norm(a):
    s = np.float32(1.0) / np.sqrt(np.maximum(_dot(a, a), np.float32(1e-30)))
    return a * s[..., None]


def render_record(rec: np.ndarray, P: Dict) -> np.ndarray:
    """uint8 image [H, W, 3] of one render record; float32 arithmetic in the order of k_render_pixels."""
    f = np.float32
    rec = np.asarray(rec, dtype=np.float32)
    W, H = P["W"], P["H"]
    o, X, Y, Z = rec[0:3], rec[3:6], rec[6:9], rec[9:12]
    nprim = int(rec[12:13].view(np.int32)[0])
    u = (np.arange(W, dtype=np.float32) + f(0.5))[None, :].repeat(H, 0).reshape(-1)
    v = (np.arange(H, dtype=np.float32) + f(0.5))[:, None].repeat(W, 1).reshape(-1)
    inv = f(1.0) / P["focal"]
    dx, dy = (u - f(0.5) * f(W)) * inv, -(v - f(0.5) * f(H)) * inv
    d = _norm(X[None, :] * dx[:, None] + Y[None, :] * dy[:, None] - Z[None, :])
    n = d.shape[0]
    tbest = np.full(n, 3.0e38, dtype=np.float32)
    mat = np.full(n, -1, dtype=np.int32)
    nrm = np.zeros((n, 3), dtype=np.float32)
    nrm[:, 2] = 1
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if o[2] > P["tab_z"]:
            m = d[:, 2] < 0
            tbest[m] = ((P["tab_z"] - o[2]) / d[m, 2]).astype(np.float32)
            mat[m] = MAT_TABLE
        for i in range(nprim):
            p = rec[REC_HDR + PRIM_FLOATS * i: REC_HDR + PRIM_FLOATS * (i + 1)]
            tm = int(p[0:1].view(np.int32)[0])
            kind, pm = tm & 15, tm >> 4
            if kind == BOX:
                R = p[7:16].reshape(3, 3)
                oc = o - p[1:4]
                ol = np.array([R[0, k] * oc[0] + R[1, k] * oc[1] + R[2, k] * oc[2] for k in range(3)], dtype=np.float32)
                dl = np.stack([R[0, k] * d[:, 0] + R[1, k] * d[:, 1] + R[2, k] * d[:, 2] for k in range(3)], axis=1)
                tn = np.full(n, -3.0e38, dtype=np.float32)
                tf = np.full(n, 3.0e38, dtype=np.float32)
                ax = np.zeros(n, dtype=np.int32)
                for k in range(3):
                    dk = np.where(np.abs(dl[:, k]) < f(1e-12), np.where(dl[:, k] < 0, f(-1e-12), f(1e-12)), dl[:, k]).astype(np.float32)
                    iv = f(1.0) / dk
                    t1 = (-ol[k]) * iv - np.abs(iv) * p[4 + k]
                    t2 = (-ol[k]) * iv + np.abs(iv) * p[4 + k]
                    upd = t1 > tn
                    tn = np.where(upd, t1, tn)
                    ax = np.where(upd, k, ax)
                    tf = np.minimum(tf, t2)
                hit = (tn <= tf) & (tn > 0) & (tn < tbest)
                sg = np.where(dl[np.arange(n), ax] > 0, f(-1), f(1))
                nn = sg[:, None] * R.T[ax]                    # column ax of R
                tbest = np.where(hit, tn, tbest)
                nrm = np.where(hit[:, None], nn, nrm).astype(np.float32)
                mat = np.where(hit, pm, mat)
                continue
            r = p[4]
            oa = o - p[1:4]
            if kind == SPHERE:
                b = d @ oa
                c = _dot(oa, oa) - r * r
                h = b * b - c
                t = -b - np.sqrt(np.maximum(h, 0))
                hit = (h > 0) & (t > 0) & (t < tbest)
                nn = (oa[None, :] + t[:, None] * d) * (f(1.0) / r)
            else:
                ba = p[5:8] - p[1:4]
                baba, baoa, oaoa = _dot(ba, ba), _dot(ba, oa), _dot(oa, oa)
                bard, rdoa = d @ ba, d @ oa
                a = baba - bard * bard
                b = baba * rdoa - baoa * bard
                c = baba * oaoa - baoa * baoa - r * r * baba
                h = b * b - a * c
                t = (-b - np.sqrt(np.maximum(h, 0))) / a
                y = baoa + t * bard
                hh = y / baba
                body = (y > 0) & (y < baba)
                lo = y <= 0
                oc = np.where(lo[:, None], oa[None, :], (o - p[5:8])[None, :]).astype(np.float32)
                b2 = _dot(d, oc)
                c2 = _dot(oc, oc) - r * r
                h2 = b2 * b2 - c2
                tcap = -b2 - np.sqrt(np.maximum(h2, 0))
                ok = (h >= 0) & (body | (h2 > 0))
                t = np.where(body, t, tcap).astype(np.float32)
                hh = np.where(body, hh, np.where(lo, f(0), f(1))).astype(np.float32)
                hit = ok & (t > 0) & (t < tbest)
                nn = (oa[None, :] + t[:, None] * d - hh[:, None] * ba[None, :]) * (f(1.0) / r)
            tbest = np.where(hit, t, tbest).astype(np.float32)
            nrm = np.where(hit[:, None], nn, nrm).astype(np.float32)
            mat = np.where(hit, pm, mat)
    V = -d
    nv = np.maximum(_dot(nrm, V), f(0))
    s = nv.copy()
    for _ in range(P["squarings"]):
        s = s * s
    s = np.where(nv > P["spec_cut"], s, f(0))           # highlight terms below 1e-6 are dropped (as the kernel does)
    dif = P["ambient"][None, :] + P["head_diffuse"][None, :] * nv[:, None]
    spc = P["head_specular"][None, :] * s[:, None]
    for l in range(len(P["ldir"])):
        L = P["ldir"][l]
        nl = nrm @ L
        Hh = _norm(L[None, :] + V)
        nh = _dot(nrm, Hh)
        sh = np.maximum(nh, f(0))
        for _ in range(P["squarings"]):
            sh = sh * sh
        sh = np.where(nh > P["spec_cut"], sh, f(0))
        on = nl > 0
        dif = dif + np.where(on[:, None], P["ldiffuse"][l][None, :] * nl[:, None], f(0))
        spc = spc + np.where(on[:, None], P["lspecular"][l][None, :] * sh[:, None], f(0))
    safe = np.maximum(mat, 0)
    rgb = np.minimum(P["mat"][safe] * dif + P["mat_specular"] * spc, f(1.0)).astype(np.float32)
    rgb = np.where((mat >= 0)[:, None], rgb, f(0))
    return (rgb * f(255.0) + f(0.5)).astype(np.int32).astype(np.uint8).reshape(H, W, 3)


def render(flat: Dict, xpos, xquat, cam_name: str, width: int, height: int) -> np.ndarray:
    return render_record(scene_record(flat, xpos, xquat, cam_name), params(flat, cam_name, width, height))
