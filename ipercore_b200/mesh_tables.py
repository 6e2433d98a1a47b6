"""Constant per-mesh tables used by the renderer seam, built from the user's asset files
(assets/configs/pose3d/mapper_uv.txt, mapper_fim_enc.txt, *.json).  One-time host-side numpy code mirroring the
table semantics of iPERCore/tools/utils/geometry/mesh.py (load_obj :50-106, get_f2vts :246-271, create_mapping
:477-540, find_part_k_nearest_faces :298-320, create_uvsampler :185-224); the results are checked against tables
produced by the reference itself (tests/golden/smpl_template.npz)."""
import json

import numpy as np


def load_obj(path):
    """Wavefront OBJ with `v`, `vt`, `f a/b/c` records -> dict(vertices, vts, faces, faces_vts) (0-based)."""
    v, vt, f, fvt = [], [], [], []
    with open(path, "r") as fp:
        for line in fp:
            s = line.split()
            if not s:
                continue
            if s[0] == "v":
                v.append([float(t) for t in s[1:4]])
            elif s[0] == "vt":
                vt.append([float(t) for t in s[1:3]])
            elif s[0] == "f":
                parts = [p.split("/") for p in s[1:4]]
                f.append([int(p[0]) - 1 for p in parts])
                if len(parts[0]) > 1:
                    fvt.append([int(p[1]) - 1 for p in parts])
    return dict(vertices=np.asarray(v, np.float32), vts=np.asarray(vt, np.float32),
                faces=np.asarray(f, np.int32), faces_vts=np.asarray(fvt, np.int32))


def face_uv_corners(obj, z):
    """(F,3,3): per-face UV corners mapped to [-1,1] with v flipped, third coordinate = z (mesh.py get_f2vts)."""
    vts = obj["vts"].astype(np.float32).copy()
    vts[:, 1] = 1 - vts[:, 1]
    vts = vts * 2 - 1
    vts = np.concatenate([vts, np.zeros((vts.shape[0], 1), np.float32) + z], axis=-1)
    return vts[obj["faces_vts"]]


def barycenter(f2vts):
    v2 = f2vts[:, 2]
    return v2 + 0.5 * (f2vts[:, 0] - v2) + 0.5 * (f2vts[:, 1] - v2)


def uv_seg_mapping(obj):
    """map_fn for map_name="uv_seg": per-face UV barycentre (u, v, 0) + background row (0, 0, 1)."""
    fbc = barycenter(face_uv_corners(obj, z=0))
    return np.concatenate([fbc, np.array([[0, 0, 1]], np.float32)], 0).astype(np.float32)


def face_flag_mapping(nf, json_path):
    """map_fn for "front"/"head"/"facial": 1 for the listed faces, background row 0."""
    m = np.zeros((nf + 1, 1), np.float32)
    with open(json_path, "r") as fp:
        m[json.load(fp)["face"]] = 1.0
    return m


def part_face_ids(nf, part_json):
    with open(part_json, "r") as fp:
        data = json.load(fp)
    parts = {k: data[k]["face"] for k in sorted(data)}
    assert len(set().union(*[set(v) for v in parts.values()])) == nf
    return parts


def part_k_nearest_faces(f2vts, parts, k):
    """For every face the k UV-nearest faces of the same body part (squared distance between barycentres)."""
    fbc = barycenter(f2vts)                       # float32 products accumulated into a float64 table, as upstream
    out = np.empty((fbc.shape[0], k), np.int64)
    for ids in parts.values():
        ids = np.asarray(ids, np.int64)
        p = fbc[ids]
        sq = (p ** 2).sum(1)
        d = np.zeros((len(ids), len(ids)))
        d += sq[None, :]
        d += sq[:, None]
        d -= 2 * p.dot(p.T)
        out[ids] = ids[np.argsort(d, axis=-1)[:, :k]]
    return out


def uv_sampler(obj, tex_size):
    """(F, T*T, 2) texture sample positions in [-1,1] (mesh.py create_uvsampler)."""
    ab = np.arange(tex_size, dtype=np.float32) / (tex_size - 1)
    coords = np.stack([(a, b) for a in ab for b in ab])
    vts = obj["vts"].astype(np.float32).copy()
    vts[:, 1] = 1 - vts[:, 1]
    f = vts[obj["faces_vts"]]
    v2 = f[:, 2]
    samples = np.dstack([f[:, 0] - v2, f[:, 1] - v2]).dot(coords.T) + v2.reshape(-1, 2, 1)
    return (np.clip(samples, 0.0, 1.0).transpose(0, 2, 1) * 2 - 1).astype(np.float32)
