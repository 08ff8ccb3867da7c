"""Seeded synthetic stand-ins for the reference datasets (no datasets exist offline).

Output tuple layout = the reference loaders' (datasets/dataset_kittisf.py:119-122):
    pcs (T,N,3) f32, segms (T,N) i32, flows (T,N,3) f32, valids (T,N) f32,  T = 2 (4 with augmentation)
A batch stacks samples on a leading axis exactly like torch's default collate.

`kittisf_like`: a KITTI-SF-shaped road scene -- ~half the points on a ground
plane, 5-8 box objects with surface points, the rest on facade planes, camera frame, depth < 35 m,
decentralised (dataset_kittisf.py:97-99); flow = ego rigid motion + per-object rigid motion + noise;
frame 2 = frame-1 geometry moved by the flow and re-sampled, with its own backward flow.
`augment` follows utils/data_util.py:140-195 (scale U[.95,1.05]^3, yaw U[-180,180] deg about y,
shift U +-[1,.1,1]); two independent views -> T = 4 ordered [v1f1, v1f2, v2f1, v2f2].
"""
import numpy as np
import torch


def _rot(axis, ang):
    c, s = np.cos(ang), np.sin(ang)
    if axis == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    if axis == "z":
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _box_surface(rng, n, size):
    """n points on the surface of an axis-aligned box of the given size, centred at the origin."""
    face = rng.integers(0, 6, size=n)
    p = rng.uniform(-0.5, 0.5, size=(n, 3))
    ax = face % 3
    p[np.arange(n), ax] = np.where(face < 3, -0.5, 0.5)
    return p * size


def kittisf_like_scene(rng, n_point=8192):
    """-> pcs (2,N,3), segms (2,N), flows (2,N,3) float64/int before casting."""
    n_obj = int(rng.integers(5, 9))
    n_ground = n_point // 2
    n_per_obj = (n_point // 4) // n_obj
    n_facade = n_point - n_ground - n_per_obj * n_obj

    def sample_static():
        ground = np.stack([rng.uniform(-25, 25, n_ground), -1.6 + rng.normal(0, 0.05, n_ground),
                           rng.uniform(5, 35, n_ground)], 1)
        side = rng.integers(0, 2, n_facade) * 2 - 1
        facade = np.stack([side * (12 + rng.normal(0, 0.05, n_facade)), rng.uniform(-1.6, 4, n_facade),
                           rng.uniform(5, 35, n_facade)], 1)
        return ground, facade

    obj_size = np.stack([rng.uniform(3.5, 4.5, n_obj), rng.uniform(1.4, 1.8, n_obj), rng.uniform(1.6, 2.0, n_obj)], 1)
    obj_pos = np.stack([rng.uniform(-9, 9, n_obj), -1.6 + obj_size[:, 1] / 2, rng.uniform(8, 32, n_obj)], 1)
    obj_yaw = rng.uniform(-np.pi, np.pi, n_obj)
    # motions between the two frames, expressed in the camera frame of frame 1
    ego_R = _rot("y", np.deg2rad(rng.uniform(-2, 2)))
    ego_t = np.array([rng.uniform(-0.1, 0.1), 0.0, -rng.uniform(0.3, 1.0)])
    obj_R = [_rot("y", np.deg2rad(rng.uniform(-5, 5))) for _ in range(n_obj)]
    obj_t = [np.array([rng.uniform(-0.3, 0.3), 0.0, rng.uniform(-2, 2)]) for _ in range(n_obj)]

    def build_frame(moved):
        ground, facade = sample_static()
        pts, seg = [ground, facade], [np.zeros(len(ground), int), np.zeros(len(facade), int)]
        for o in range(n_obj):
            local = _box_surface(rng, n_per_obj, obj_size[o]) @ _rot("y", obj_yaw[o]).T + obj_pos[o]
            pts.append(local)
            seg.append(np.full(n_per_obj, o + 1))
        pts, seg = np.concatenate(pts), np.concatenate(seg)
        # forward motion of every point (frame 1 -> frame 2) in frame-1 coordinates
        fwd = pts @ ego_R.T + ego_t
        for o in range(n_obj):
            sel = seg == o + 1
            centre = obj_pos[o]
            fwd[sel] = ((pts[sel] - centre) @ obj_R[o].T + centre + obj_t[o]) @ ego_R.T + ego_t
        if not moved:
            return pts, seg, fwd - pts
        return fwd, seg, pts - fwd          # frame 2: geometry after the motion, flow points back

    pc1, seg1, flow1 = build_frame(False)
    pc2, seg2, flow2 = build_frame(True)
    noise = lambda f: f + rng.normal(0, 0.02, f.shape)
    pcs = np.stack([pc1, pc2])
    pcs = pcs - pcs.mean(1).mean(0)                        # decentralize
    perm1, perm2 = rng.permutation(n_point), rng.permutation(n_point)
    pcs = np.stack([pcs[0][perm1], pcs[1][perm2]])
    segms = np.stack([seg1[perm1], seg2[perm2]])
    flows = np.stack([noise(flow1)[perm1], noise(flow2)[perm2]])
    return pcs, segms, flows


def augment(rng, pcs, flows, n_view=2):
    """utils/data_util.py:140-195 with the KITTI-SF arguments (config/seg/kittisf/kittisf_unsup.yaml:9-13)."""
    out_p, out_f = [], []
    for _ in range(n_view):
        rot = _rot("y", np.deg2rad(rng.uniform(-180, 180)))
        scale = rng.uniform(0.95, 1.05, 3)
        shift = rng.uniform(-np.array([1, 0.1, 1]), np.array([1, 0.1, 1]))
        for t in range(2):
            out_p.append(scale * (pcs[t] @ rot.T) + shift)
            out_f.append(scale * (flows[t] @ rot.T))
    return np.stack(out_p), np.stack(out_f)


def fps_order(pcs, fps_fn):
    """Store clouds in furthest-point-sampling order like the reference's down-sampled KITTI-SF
    (data_prepare/kittisf/downsample_kittisf.py:49-52).  pcs (...,N,3) tensor; fps_fn(xyz, N) -> idx."""
    flat = pcs.reshape(-1, pcs.shape[-2], 3).contiguous()
    order = fps_fn(flat, flat.shape[1]).long()
    return order.reshape(*pcs.shape[:-1])


def make_batch(seed, batch_size, n_point=8192, aug=True, fps_fn=None, device="cpu"):
    """-> pcs (b,T,N,3) f32, segms (b,T,N) i32, flows (b,T,N,3) f32, valids (b,T,N) f32 on CPU
    (pinned when CUDA exists) -- what the reference DataLoader hands to Trainer._train_it."""
    rng = np.random.default_rng(seed)
    P, S, F_ = [], [], []
    for _ in range(batch_size):
        pcs, segms, flows = kittisf_like_scene(rng, n_point)
        if fps_fn is not None:
            t = torch.from_numpy(pcs.astype(np.float32)).to(device)
            order = fps_order(t, fps_fn).cpu().numpy()
            pcs = np.stack([pcs[i][order[i]] for i in range(2)])
            segms = np.stack([segms[i][order[i]] for i in range(2)])
            flows = np.stack([flows[i][order[i]] for i in range(2)])
        if aug:
            pcs, flows = augment(rng, pcs, flows)
            segms = np.concatenate([segms, segms], 0)
        P.append(pcs); S.append(segms); F_.append(flows)
    pcs = torch.from_numpy(np.stack(P).astype(np.float32))
    segms = torch.from_numpy(np.stack(S).astype(np.int32))
    flows = torch.from_numpy(np.stack(F_).astype(np.float32))
    valids = torch.ones(segms.shape, dtype=torch.float32)
    if torch.cuda.is_available():
        pcs, segms, flows, valids = (x.pin_memory() for x in (pcs, segms, flows, valids))
    return pcs, segms, flows, valids
