"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path (siu3r_b200/).

NumPy (float32, op-by-op) restatement of the N-channel splatting the reference obtains from
    gsplat.rasterization(means, quats=None, scales=None, covars, opacities, colors[N,C], viewmats, Ks, width, height,
                         sh_degree=None, near_plane, far_plane)            (/root/reference/src/models/gaussian_renderer.py:92-106)
PARITY UNPINNED: gsplat (1.5.2 @ 961678f, uv.lock:757-759) is an un-vendored CUDA dependency that is absent from /root/reference and
from this image; the reference holds no vectors for it.  This follows the published classic-mode algorithm of that release:
  projection  : camera-space mean / covariance, pinhole EWA Jacobian with the (W-cx)/fx + 0.3 tan(fov/2) clamp, eps2d = 0.3 blur
  extents     : opacity-aware, min(3.33, sqrt(2 ln(o / (1/255)))) sigma per axis, capped by the major-axis radius; 16x16 tiles
  ordering    : per tile by (depth bits, Gaussian id) -- a stable radix sort of (tile << 32 | depth) keys
  blend       : pixel centres at +0.5; sigma = d^T conic d / 2; alpha = min(0.999, o exp(-sigma)); skipped if sigma < 0 or alpha < 1/255;
                a pixel stops BEFORE the Gaussian that would take T to <= 1e-4; out = sum feature * alpha * T, alpha_out = 1 - T.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def project(means, covars, opac, viewmat, fx, fy, cx, cy, W, H, near, far):
    """-> dict(valid [G] bool, xy [G,2], conic [G,3], depth [G], radii [G,2] int32, rect [G,4] int32 (x0, y0, x1, y1 in tiles))."""
    means, covars, opac, V = means.astype(F), covars.astype(F), opac.astype(F), viewmat.astype(F)
    fx, fy, cx, cy = F(fx), F(fy), F(cx), F(cy)
    G = means.shape[0]
    px, py, pz = means[:, 0], means[:, 1], means[:, 2]
    x = ((V[0, 0] * px + V[0, 1] * py) + V[0, 2] * pz) + V[0, 3]
    y = ((V[1, 0] * px + V[1, 1] * py) + V[1, 2] * pz) + V[1, 3]
    z = ((V[2, 0] * px + V[2, 1] * py) + V[2, 2] * pz) + V[2, 3]
    valid = ~((z < F(near)) | (z > F(far)))
    S = covars.reshape(G, 3, 3)
    RS = np.empty((G, 3, 3), F)
    Cc = np.empty((G, 3, 3), F)
    for r in range(3):
        for q in range(3):
            RS[:, r, q] = (V[r, 0] * S[:, 0, q] + V[r, 1] * S[:, 1, q]) + V[r, 2] * S[:, 2, q]
    for r in range(3):
        for q in range(3):
            Cc[:, r, q] = (RS[:, r, 0] * V[q, 0] + RS[:, r, 1] * V[q, 1]) + RS[:, r, 2] * V[q, 2]
    tanx, tany = (F(0.5) * F(W)) / fx, (F(0.5) * F(H)) / fy
    lxp, lxn = (F(W) - cx) / fx + F(0.3) * tanx, cx / fx + F(0.3) * tanx
    lyp, lyn = (F(H) - cy) / fy + F(0.3) * tany, cy / fy + F(0.3) * tany
    with np.errstate(all="ignore"):
        rz = F(1.0) / z
        rz2 = rz * rz
        tx = z * np.minimum(lxp, np.maximum(-lxn, x * rz))
        ty = z * np.minimum(lyp, np.maximum(-lyn, y * rz))
        ja, jb, jc, jd = fx * rz, -((fx * tx) * rz2), fy * rz, -((fy * ty) * rz2)
        JC0 = [ja * Cc[:, 0, q] + jb * Cc[:, 2, q] for q in range(3)]
        JC1 = [jc * Cc[:, 1, q] + jd * Cc[:, 2, q] for q in range(3)]
        c00 = JC0[0] * ja + JC0[2] * jb
        c01 = JC0[1] * jc + JC0[2] * jd
        c11 = JC1[1] * jc + JC1[2] * jd
        mx, my = (fx * x) * rz + cx, (fy * y) * rz + cy
        c00, c11 = c00 + F(0.3), c11 + F(0.3)
        det = c00 * c11 - c01 * c01
        valid &= det > 0
        thr = F(1.0) / F(255.0)
        valid &= ~(opac < thr)
        ext = np.minimum(F(3.33), np.sqrt(F(2.0) * np.log(opac / thr)))
        bb = F(0.5) * (c00 + c11)
        v1 = bb + np.sqrt(np.maximum(F(0.01), bb * bb - det))
        r1 = ext * np.sqrt(v1)
        rx = np.ceil(np.minimum(ext * np.sqrt(c00), r1))
        ry = np.ceil(np.minimum(ext * np.sqrt(c11), r1))
        valid &= ~((rx <= 0) & (ry <= 0))
        valid &= ~((mx + rx <= 0) | (mx - rx >= F(W)) | (my + ry <= 0) | (my - ry >= F(H)))
        gx, gy = (W + 15) // 16, (H + 15) // 16
        tsx, tsy, trx, try_ = mx / F(16), my / F(16), rx / F(16), ry / F(16)
        f2i = lambda a: np.nan_to_num(a, nan=0.0, posinf=1e9, neginf=-1e9).astype(np.int64)
        x0 = np.minimum(gx, np.maximum(0, f2i(np.floor(tsx - trx))))
        x1 = np.minimum(gx, np.maximum(0, f2i(np.ceil(tsx + trx))))
        y0 = np.minimum(gy, np.maximum(0, f2i(np.floor(tsy - try_))))
        y1 = np.minimum(gy, np.maximum(0, f2i(np.ceil(tsy + try_))))
        valid &= (x1 - x0) * (y1 - y0) > 0
        inv = F(1.0) / det
        conic = np.stack((c11 * inv, (-c01) * inv, c00 * inv), -1)
    radii = np.where(valid[:, None], np.stack((rx, ry), -1), 0).astype(np.int32)
    return dict(valid=valid, xy=np.stack((mx, my), -1), conic=conic, depth=z, radii=radii, rect=np.stack((x0, y0, x1, y1), -1).astype(np.int32))


def rasterize(means, covars, opac, feats, viewmat, fx, fy, cx, cy, W, H, near, far):
    """-> (features [H, W, C], alpha [H, W], projection dict)."""
    pr = project(means, covars, opac, viewmat, fx, fy, cx, cy, W, H, near, far)
    feats, opac = feats.astype(F), opac.astype(F)
    C = feats.shape[1]
    out = np.zeros((H, W, C), F)
    alpha_out = np.zeros((H, W), F)
    ids_all = np.nonzero(pr["valid"])[0]
    dbits = pr["depth"].astype(F).view(np.uint32)
    rect = pr["rect"]
    for ty in range((H + 15) // 16):
        for tx in range((W + 15) // 16):
            m = (rect[ids_all, 0] <= tx) & (tx < rect[ids_all, 2]) & (rect[ids_all, 1] <= ty) & (ty < rect[ids_all, 3])
            ids = ids_all[m]
            ids = ids[np.lexsort((ids, dbits[ids]))]        # stable by depth bits, ties by id
            ys, xs = np.meshgrid(np.arange(ty * 16, min(H, ty * 16 + 16)), np.arange(tx * 16, min(W, tx * 16 + 16)), indexing="ij")
            pxf, pyf = xs.astype(F) + F(0.5), ys.astype(F) + F(0.5)
            T = np.ones(pxf.shape, F)
            done = np.zeros(pxf.shape, bool)
            acc = np.zeros(pxf.shape + (C,), F)
            for g in ids:
                if done.all():
                    break
                dx, dy = pr["xy"][g, 0] - pxf, pr["xy"][g, 1] - pyf
                a, b, c = pr["conic"][g]
                sigma = F(0.5) * (a * dx * dx + c * dy * dy) + b * dx * dy
                al = np.minimum(F(0.999), opac[g] * np.exp(-sigma))
                use = ~done & ~(sigma < 0) & ~(al < F(1.0) / F(255.0))
                nT = T * (F(1.0) - al)
                stop = use & (nT <= F(1e-4))
                done |= stop
                use &= ~stop
                vis = np.where(use, al * T, F(0))
                acc += vis[..., None] * feats[g]
                T = np.where(use, nT, T)
            out[ys, xs] = acc
            alpha_out[ys, xs] = F(1.0) - T
    return out, alpha_out, pr
