"""Synthetic keyframe SEQUENCE for BASELINE.json configs[3] ("KITTI 1241x376 preset=0 tracker+BA loop on synthetic sequence"):
frames rendered along a smooth fly-through of the procedural scene of ldso_b200.synth (1232x368
after the KITTI crop, 5 pyramid levels, KITTI intrinsics), every `kf_every`-th frame a keyframe with `pts_per_kf` active points, a
sliding window of the last `window` keyframes. Inputs only: the loop itself (tracker on every frame, photometric BA on every
keyframe) is driven through the C ABI by bench.py / tests."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import synth

KITTI_K = np.array([718.856, 718.856, 607.1928, 185.2157])      # examples/Kitti/Kitti00-02.txt:1-4, cropped to 1232x368


@dataclasses.dataclass
class Sequence:
    w: int
    h: int
    levels: int
    K: np.ndarray
    n_frames: int
    kf_every: int
    window: int
    Rcw: np.ndarray          # (n,3,3) ground-truth worldToCam
    tcw: np.ndarray          # (n,3)
    aff: np.ndarray          # (n,2) a, b of every frame
    images: list             # n x (h,w) f32 raw intensity (what FrameHessian::makeImages receives)
    kf_points: dict          # keyframe frame index -> dict(u, v, idepth_zero, idepth, color8, weights8)

    def pose(self, k):
        return self.Rcw[k], self.tcw[k]

    def rel_pose(self, ref, new):
        """refToNew = T_new * T_ref^-1 (what trackNewestCoarse estimates)."""
        R = self.Rcw[new] @ self.Rcw[ref].T
        return R, self.tcw[new] - R @ self.tcw[ref]

    def window_kfs(self, k):
        """Frame indices of the keyframes in the sliding window when frame k (a keyframe) has been inserted."""
        kfs = list(range(0, k + 1, self.kf_every))
        return kfs[-self.window:]


def make_sequence(n_frames=50, w=1232, h=368, K=None, kf_every=5, window=8, pts_per_kf=250, seed=42, baseline=0.06,
                  outlier_frac=0.05) -> Sequence:
    rng = np.random.default_rng(seed)
    K = np.asarray(KITTI_K if K is None else K, np.float64)
    levels = synth.pyr_levels_for(w, h)
    tex = synth.Texture(1234)
    Rcw = np.zeros((n_frames, 3, 3)); tcw = np.zeros((n_frames, 3)); aff = np.zeros((n_frames, 2))
    images = []
    for k in range(n_frames):
        s = k / float(kf_every)             # keyframes are spaced like synth.make_window's
        Rwc = synth.so3_exp(0.01 * s * np.array([0.3, 1.0, 0.2]))
        twc = np.array([baseline * s, 0.01 * np.sin(s), 0.02 * s / 0.06 * baseline])
        Rcw[k] = Rwc.T
        tcw[k] = -Rwc.T @ twc
        if k > 0:
            aff[k] = [rng.uniform(-0.03, 0.03), rng.uniform(-3.0, 3.0)]
        images.append(np.ascontiguousarray(synth.render(tex, Rcw[k], tcw[k], K, w, h, aff[k, 0], aff[k, 1]), np.float32))
    kf_points = {}
    for k in range(0, n_frames, kf_every):
        pyr0 = synth.make_images(images[k], 1)[0]
        u = rng.integers(20, w - 20, pts_per_kf).astype(np.float64)
        v = rng.integers(20, h - 20, pts_per_kf).astype(np.float64)
        _, _, depth, _ = synth.scene_depth(Rcw[k], tcw[k], K, u, v)
        id_true = 1.0 / depth
        id_zero = id_true * (1.0 + rng.normal(0.0, 0.01, pts_per_kf))
        is_out = rng.uniform(0, 1, pts_per_kf) < outlier_frac
        id_zero = np.where(is_out, id_zero * rng.uniform(0.3, 3.0, pts_per_kf), id_zero)
        id_cur = np.maximum(id_zero + rng.normal(0.0, 5e-3, pts_per_kf), 1e-3)
        c8 = np.zeros((pts_per_kf, 8), np.float32); w8 = np.zeros((pts_per_kf, 8), np.float32)
        for q in range(8):
            c, gx, gy = synth.sample_bilin(pyr0, u + synth.PATTERN[q, 0], v + synth.PATTERN[q, 1])
            c8[:, q] = c
            w8[:, q] = np.sqrt(np.float32(synth.OUTLIER_TH_SUM_COMPONENT) / (np.float32(synth.OUTLIER_TH_SUM_COMPONENT) + (gx * gx + gy * gy)))
        kf_points[k] = dict(u=u.astype(np.float32), v=v.astype(np.float32), idepth_zero=id_zero.astype(np.float32),
                            idepth=id_cur.astype(np.float32), color=c8, weights=w8)
    return Sequence(w=w, h=h, levels=levels, K=K, n_frames=n_frames, kf_every=kf_every, window=window, Rcw=Rcw, tcw=tcw, aff=aff,
                    images=images, kf_points=kf_points)


def window_arrays(seq: Sequence, kfs, seed=0):
    """Flattened window (ldso_b200_set_frames / set_window inputs) for the keyframes `kfs` (frame indices, oldest first): evaluation
    point = ground truth, state = evaluation point + a small perturbation (what a tracked / previously optimised pose looks like),
    every point observed in every other keyframe of the window."""
    rng = np.random.default_rng(seed + 7919 * kfs[-1])
    nF = len(kfs)
    state_zero = np.zeros((nF, 10))
    state_zero[:, 6] = seq.aff[kfs, 0] / synth.SCALE_A
    state_zero[:, 7] = seq.aff[kfs, 1] / synth.SCALE_B
    state = state_zero.copy()
    d_pose = rng.normal(0.0, 2e-3, (nF, 6)); d_ab = rng.normal(0.0, 1e-3, (nF, 2))
    if kfs[0] == 0:
        d_pose[0] = 0.0; d_ab[0] = 0.0
    state[:, 0:3] += d_pose[:, 0:3] / synth.SCALE_XI_TRANS
    state[:, 3:6] += d_pose[:, 3:6] / synth.SCALE_XI_ROT
    state[:, 6] += d_ab[:, 0] / synth.SCALE_A
    state[:, 7] += d_ab[:, 1] / synth.SCALE_B
    P = [seq.kf_points[k] for k in kfs]
    n_per = [len(p["u"]) for p in P]
    pt_host = np.concatenate([np.full(n, i, np.int32) for i, n in enumerate(n_per)])
    cat = lambda key: np.concatenate([p[key] for p in P])
    targets = [np.array([t for t in range(nF) if t != h], np.int32) for h in range(nF)]
    res_target = np.concatenate([np.tile(targets[h], n_per[h]) for h in range(nF)]) if nF > 1 else np.zeros(0, np.int32)
    res_begin = np.zeros(len(pt_host) + 1, np.int32)
    res_begin[1:] = np.cumsum(np.concatenate([np.full(n_per[h], nF - 1, np.int32) for h in range(nF)]))
    return dict(Rcw=seq.Rcw[kfs], tcw=seq.tcw[kfs], state_zero=state_zero, state=state, ab_exposure=np.ones(nF, np.float32),
                frame_id=np.asarray(kfs, np.int32) // seq.kf_every, pt_host=pt_host, pt_u=cat("u"), pt_v=cat("v"),
                pt_idepth=cat("idepth"), pt_idepth_zero=cat("idepth_zero"), pt_has_prior=np.zeros(len(pt_host), np.uint8),
                pt_color=cat("color"), pt_weights=cat("weights"), res_begin=res_begin, res_target=res_target)
