"""CPU: host-side geometry code of stage II -- the PLY reader that replaces open3d.io.read_triangle_mesh
(network/renderer.py:675) and the BVH builder behind nero_bvh_build_host (a host function of the C-ABI library, callable
without a GPU).  The tree is validated structurally and by a reference traversal in numpy against the oracle's
exhaustive tracer."""
import struct

import numpy as np
import torch

import nero_oracle as O
import nero_oracle_mat as OM

NODE = np.dtype([('lo', '<f4', (3,)), ('a', '<i4'), ('hi', '<f4', (3,)), ('count', '<i4')])


def test_read_ply_ascii_and_binary(tmp_path):
    from nero_b200.material import read_ply
    verts, tris = OM.test_scene(1)
    a = tmp_path / 'a.ply'
    with open(a, 'w') as f:
        f.write(f'ply\nformat ascii 1.0\ncomment made by a test\nelement vertex {len(verts)}\nproperty float x\nproperty float y\n'
                f'property float z\nelement face {len(tris)}\nproperty list uchar int vertex_indices\nend_header\n')
        for v in verts:
            f.write('%.9g %.9g %.9g\n' % tuple(v))
        for t in tris:
            f.write('3 %d %d %d\n' % tuple(t))
    b = tmp_path / 'b.ply'
    with open(b, 'wb') as f:
        f.write((f'ply\nformat binary_little_endian 1.0\nelement vertex {len(verts)}\nproperty double x\nproperty double y\n'
                 f'property double z\nproperty uchar red\nelement face {len(tris)}\nproperty list uchar uint vertex_indices\n'
                 f'end_header\n').encode())
        for v in verts:
            f.write(struct.pack('<dddB', *[float(x) for x in v], 7))
        for t in tris:
            f.write(struct.pack('<BIII', 3, *[int(x) for x in t]))
    for path in (a, b):
        v, t = read_ply(str(path))
        assert v.dtype == np.float32 and t.dtype == np.int32
        np.testing.assert_allclose(v, verts, rtol=1e-7)
        np.testing.assert_array_equal(t, tris)


def test_bvh_build_structure_and_traversal():
    from nero_b200 import ops
    verts, tris = OM.test_scene(3)
    T = tris.shape[0]
    nodes_u8, tri, ids = ops.bvh_build(verts, tris)
    nodes = nodes_u8.view(NODE).reshape(-1)
    assert sorted(ids.tolist()) == list(range(T)), 'every triangle is referenced exactly once'
    # re-ordered triangle records = (v0, e1, e2) of the original triangle ids[k]
    np.testing.assert_allclose(tri[:, 0:3], verts[tris[ids, 0]], rtol=0, atol=0)
    np.testing.assert_allclose(tri[:, 4:7], verts[tris[ids, 1]] - verts[tris[ids, 0]], rtol=0, atol=1e-7)
    # structure: leaves partition [0, T); bounds of a node contain its triangles / children; depth bounded
    covered = np.zeros(T, np.int32)
    max_depth = 0
    stack = [(0, 0)]
    while stack:
        i, depth = stack.pop()
        n = nodes[i]
        max_depth = max(max_depth, depth)
        if n['count'] > 0:
            assert n['count'] <= 4 or depth > 0
            ks = np.arange(n['a'], n['a'] + n['count'])
            covered[ks] += 1
            v0 = tri[ks, 0:3]
            pts = np.concatenate([v0, v0 + tri[ks, 4:7], v0 + tri[ks, 8:11]])
            assert (pts >= n['lo'] - 1e-6).all() and (pts <= n['hi'] + 1e-6).all()
        else:
            left, right = i + 1, int(n['a'])
            for c in (left, right):
                assert (nodes[c]['lo'] >= n['lo'] - 1e-6).all() and (nodes[c]['hi'] <= n['hi'] + 1e-6).all()
                stack.append((c, depth + 1))
    assert (covered == 1).all() and max_depth < 44, max_depth

    # traversal of that tree (slab test, depth-first, left child = index + 1) reproduces the exhaustive tracer
    def walk(o, d):
        best, st = 10.0, [0]
        inv = 1.0 / np.where(d == 0, 1e-30, d)
        while st:
            i = st.pop()
            n = nodes[i]
            t0, t1 = (n['lo'] - o) * inv, (n['hi'] - o) * inv
            tmin, tmx = np.minimum(t0, t1).max(), np.maximum(t0, t1).min()
            if tmx < max(tmin, 0.0) - 1e-6 or tmin > best:
                continue
            if n['count'] > 0:
                for k in range(n['a'], n['a'] + n['count']):
                    v0, e1, e2 = (tri[k, 0:3].astype(np.float64), tri[k, 4:7].astype(np.float64), tri[k, 8:11].astype(np.float64))
                    p = np.cross(d, e2)
                    det = e1 @ p
                    if abs(det) <= 1e-12:
                        continue
                    tv = o - v0
                    u, q = (tv @ p) / det, np.cross(tv, e1)
                    v, t = (d @ q) / det, (e2 @ q) / det
                    if u >= 0 and v >= 0 and u + v <= 1 and 0 < t < best:
                        best = t
            else:
                st.append(int(n['a']))
                st.append(i + 1)
        return best
    rays = O.synthetic_rays(150, seed=8)
    _, _, depth = OM.trace_bruteforce(verts, tris, rays['rays_o'].double(), rays['rays_d'].double())
    got = np.array([walk(o, d) for o, d in zip(rays['rays_o'].double().numpy(), rays['rays_d'].double().numpy())])
    assert (depth.numpy() < 10).sum() > 50
    np.testing.assert_allclose(got, depth.numpy(), rtol=1e-6, atol=1e-6)
