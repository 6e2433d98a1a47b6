"""CPU restatement of SMPL/SMPLH linear blend skinning (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows /root/reference/iPERCore/tools/human_digitalizer/smplx/lbs.py:137-227 (lbs), :259-271 (blend_shapes),
:321-375 (batch_rigid_transform), iPERCore/tools/utils/geometry/rotations.py:318-332,355-375 (rotvec -> quaternion ->
matrix), bodynets/batch_smplh.py:137-180 (72-dim pose + hands_mean) and bodynets/base_smpl.py:28-50 (link).
Pinned against the reference's own `lbs()` by tests/golden/make_golden.py (lbs.npz) on the synthetic model below —
the real SMPL pkl is not available offline.
"""
import numpy as np


def synthetic_smplh(seed=7, nv=6890, nj=52, nb=10, template=None):
    """SMPLH-shaped random model: random kinematic tree, sparse non-negative joint regressor and skinning weights
    (<= 4 joints per vertex, rows sum to 1), small shape / pose blend directions.  Pure function of the seed."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parents = np.array([-1] + [int(rng.integers(max(0, i - 4), i)) for i in range(1, nj)], np.int64)
    v_template = (template if template is not None else rng.uniform(-1, 1, (nv, 3))).astype(np.float32)
    shapedirs = (rng.standard_normal((nv, 3, nb)) * 0.02).astype(np.float32)
    posedirs = (rng.standard_normal(((nj - 1) * 9, nv * 3)) * 0.005).astype(np.float32)
    J_regressor = np.zeros((nj, nv), np.float32)
    for j in range(nj):
        idx = rng.choice(nv, 24, replace=False)
        w = rng.uniform(0.1, 1.0, 24); J_regressor[j, idx] = (w / w.sum()).astype(np.float32)
    lbs_weights = np.zeros((nv, nj), np.float32)
    for v in range(nv):
        idx = rng.choice(nj, 4, replace=False)
        w = rng.uniform(0.05, 1.0, 4); lbs_weights[v, idx] = (w / w.sum()).astype(np.float32)
    hands_mean = (rng.standard_normal(90) * 0.1).astype(np.float32)
    return dict(v_template=v_template, shapedirs=shapedirs, posedirs=posedirs, J_regressor=J_regressor,
                parents=parents, lbs_weights=lbs_weights, hands_mean=hands_mean)


def rotvec_to_rotmat(rv):
    """rotations.py:318-332 + :355-375, float32 throughout.  rv (N,3) -> (N,3,3)."""
    rv = rv.astype(np.float32)
    angle = np.sqrt(((rv + np.float32(1e-8)) ** 2).sum(1, keepdims=True)).astype(np.float32)
    normalized = rv / angle
    half = angle * np.float32(0.5)
    quat = np.concatenate([np.cos(half), np.sin(half) * normalized], 1).astype(np.float32)
    quat = quat / np.sqrt((quat ** 2).sum(1, keepdims=True)).astype(np.float32)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    R = np.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                  2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                  2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], 1)
    return R.reshape(-1, 3, 3).astype(np.float32)


def lbs(model, betas, full_pose, offsets=0.0, links=None):
    """betas (B,10), full_pose (B, nj*3) axis-angle -> verts (B,V,3), posed joints (B,nj,3)  [lbs.py:137-227]."""
    m = model
    B = full_pose.shape[0]; nj = m["J_regressor"].shape[0]
    v_template = m["v_template"] + offsets
    v_shaped = v_template[None] + np.einsum("bl,mkl->bmk", betas.astype(np.float32), m["shapedirs"])
    J = np.einsum("jv,bvk->bjk", m["J_regressor"], v_shaped).astype(np.float32)
    R = rotvec_to_rotmat(full_pose.reshape(-1, 3)).reshape(B, nj, 3, 3)
    pose_feature = (R[:, 1:] - np.eye(3, dtype=np.float32)).reshape(B, -1)
    v_posed = v_shaped + (pose_feature @ m["posedirs"]).reshape(B, -1, 3)
    # batch_rigid_transform (lbs.py:321-375)
    rel = J.copy(); rel[:, 1:] -= J[:, m["parents"][1:]]
    T = np.zeros((B, nj, 4, 4), np.float32); T[:, :, :3, :3] = R; T[:, :, :3, 3] = rel; T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, nj):
        chain.append(chain[m["parents"][i]] @ T[:, i])
    G = np.stack(chain, 1)
    posed_joints = G[:, :, :3, 3].copy()
    A = G.copy()
    A[:, :, :3, 3] -= np.einsum("bjik,bjk->bji", G[:, :, :3, :3], J)
    Tv = np.einsum("vj,bjrc->bvrc", m["lbs_weights"], A)
    verts = np.einsum("bvrc,bvc->bvr", Tv[:, :, :3, :3], v_posed) + Tv[:, :, :3, 3]
    verts = verts.astype(np.float32)
    if links is not None:           # base_smpl.py:28-50, 2-D form: verts[:, links[:,0]] = verts[:, links[:,1]]
        out = verts.copy(); out[:, links[:, 0]] = verts[:, links[:, 1]]; verts = out
    return verts, posed_joints.astype(np.float32)


def smplh_forward(model, beta, theta, offsets=0.0, links=None):
    """SMPLH.forward (batch_smplh.py:137-180) without hand PCA: 72-dim body pose is padded with hands_mean."""
    theta = theta.astype(np.float32)
    if theta.shape[1] == 72:
        theta = np.concatenate([theta[:, :66], np.repeat(model["hands_mean"][None], theta.shape[0], 0)], 1)
    return lbs(model, beta, theta, offsets, links)
