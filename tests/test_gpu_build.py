"""GPU parity: LBVH build (rt_blas_build / rt_tlas_build through the C ABI) against the CPU oracle.

North-star criterion 1: Morton codes and the sort permutation are BIT-EXACT.  The hierarchy and the
whole reference-format blob (boxes, sorted primitives, metadata) are compared bit-exactly as well.
"""
import numpy as np
import pytest

from dxrexperiments_b200 import scenes, types as T

pytestmark = pytest.mark.gpu


def _mesh_cases():
    rng = np.random.Generator(np.random.PCG64(5))
    dup = scenes.triangle_soup(3000, seed=3, extent=2.0, edge=0.5)            # many duplicate Morton codes
    dup.vertices["position"][: 3 * 500] = np.tile(dup.vertices["position"][:3], (500, 1))  # 500 identical triangles
    flat = scenes.quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1))       # zero-extent axis (epsilon path)
    return {
        "cornell": scenes.cornell_box(),
        "icosphere3": scenes.icosphere(3),
        "bunny4": scenes.bunny_scale(4),
        "soup5000": scenes.triangle_soup(5000, seed=42),
        "duplicates": dup,
        "flat_quad": flat,
        "one_triangle": scenes.Mesh(scenes.icosphere(0).vertices, scenes.icosphere(0).indices[:3].copy()),
        "two_triangles": scenes.Mesh(scenes.icosphere(0).vertices, scenes.icosphere(0).indices[:6].copy()),
        "soup_70k": scenes.triangle_soup(70001, seed=int(rng.integers(1 << 30))),
    }


MESHES = _mesh_cases()


# treelet pass count by build flag (FL/TreeletReorder.cpp:66-80): default 1, PREFER_FAST_BUILD 0, PREFER_FAST_TRACE 3
BUILD_FLAGS = {"default": 0, "fast_build": T.BUILD_FLAG_PREFER_FAST_BUILD, "fast_trace": T.BUILD_FLAG_PREFER_FAST_TRACE}


@pytest.mark.parametrize("variant", sorted(BUILD_FLAGS))
@pytest.mark.parametrize("name", sorted(MESHES))
def test_blas_stages_and_blob_bit_exact(name, variant, ctx, orc):
    mesh = MESHES[name]
    ref = orc.Blas.from_mesh(mesh, build_flags=BUILD_FLAGS[variant])
    acc = ctx.build_blas_from_mesh(mesh, keep_scratch=True, build_flags=BUILD_FLAGS[variant])
    ctx.sync()
    assert acc.n == ref.n == mesh.num_triangles
    np.testing.assert_array_equal(acc.stage("primitives").view(np.uint8), ref.unsorted_prims().view(np.uint8))
    np.testing.assert_array_equal(acc.stage("scene_aabb"), ref.scene_aabb())
    np.testing.assert_array_equal(acc.stage("morton_codes"), ref.morton())          # criterion 1
    np.testing.assert_array_equal(acc.stage("sorted_codes"), ref.sorted_morton())
    np.testing.assert_array_equal(acc.stage("sorted_indices"), ref.perm())           # criterion 1
    if ref.n > 1 and variant == "fast_build":
        # PREFER_FAST_BUILD emits the hierarchy and fits the boxes in ONE bottom-up kernel (k_lbvh_fit): there is no
        # hierarchy array; the child links of the blob carry the whole topology (parent links follow from them)
        h_ref, nodes = ref.hierarchy(), T.parse_blas_blob(acc.blob())["nodes"]
        n_int = ref.n - 1
        left = np.where(nodes["flags"][:n_int] & 0x80000000, 0, nodes["flags"][:n_int] & 0x00FFFFFF)
        parent = np.full(2 * ref.n - 1, 0xFFFFFFFF, np.uint32)
        parent[nodes["right"][:n_int]] = np.arange(n_int, dtype=np.uint32)
        parent[left] = np.arange(n_int, dtype=np.uint32)
        np.testing.assert_array_equal(parent[1:], h_ref["parent"][1:] & 0x7FFFFFFF)
        kids_gpu = np.sort(np.stack([left, nodes["right"][:n_int]]), axis=0)
        kids_ref = np.sort(np.stack([h_ref["left"][:n_int], h_ref["right"][:n_int]]), axis=0)
        np.testing.assert_array_equal(kids_gpu, kids_ref)
    elif ref.n > 1:
        h_gpu, h_ref = acc.stage("hierarchy"), ref.hierarchy()
        np.testing.assert_array_equal(h_gpu["left"][: ref.n - 1], h_ref["left"][: ref.n - 1])
        np.testing.assert_array_equal(h_gpu["right"][: ref.n - 1], h_ref["right"][: ref.n - 1])
        np.testing.assert_array_equal(h_gpu["parent"][1:], h_ref["parent"][1:])
    np.testing.assert_array_equal(acc.blob(), ref.blob())


def test_blas_index_formats_transform_and_multi_geometry(ctx, orc):
    """R16 / R32 / no index buffer, a 3x4 transform and two geometries in one BLAS (UT:617-776)."""
    verts, idx16 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1]], np.float32), np.array([0, 1, 2, 2, 1, 3], np.uint16)
    soup = scenes.triangle_soup(100, seed=9, extent=5.0)
    soup_pos = np.ascontiguousarray(soup.vertices["position"])
    xf = np.array([0.5, 0, 0, 1, 0, 2, 0, -3, 0.25, 0, 1, 0.5], np.float32)
    ogeoms = [dict(vertices=verts, stride=12, indices=idx16, flags=1),
              dict(vertices=soup_pos, stride=12, indices=None, transform=xf, flags=0),
              dict(vertices=verts, stride=12, indices=idx16.astype(np.uint32), flags=1)]
    ref = orc.Blas(ogeoms)
    d_verts, d_idx16, d_idx32 = ctx.upload(verts), ctx.upload(idx16), ctx.upload(idx16.astype(np.uint32))
    d_soup, d_xf = ctx.upload(soup_pos), ctx.upload(xf)
    acc = ctx.build_blas([
        dict(vertices=d_verts, vertex_count=4, stride=12, indices=d_idx16, index_count=6, index_format=16, flags=1),
        dict(vertices=d_soup, vertex_count=300, stride=12, indices=None, index_format=0, transform=d_xf, flags=0),
        dict(vertices=d_verts, vertex_count=4, stride=12, indices=d_idx32, index_count=6, index_format=32, flags=1)],
        keep_scratch=True)
    ctx.sync()
    assert acc.n == ref.n == 104
    np.testing.assert_array_equal(acc.stage("sorted_indices"), ref.perm())
    np.testing.assert_array_equal(acc.blob(), ref.blob())
    meta = T.parse_blas_blob(acc.blob())["meta"]
    assert set(meta["geom"].tolist()) == {0, 1, 2}
    assert sorted(meta["prim"][meta["geom"] == 1].tolist()) == list(range(100))


@pytest.mark.parametrize("n_inst", [1, 2, 50, 1000])
def test_tlas_blob_bit_exact(n_inst, ctx, orc):
    """TLAS over instances with random rigid transforms (UT:778-935 use 1 and 50)."""
    mesh = scenes.icosphere(2)
    xf = scenes.random_rigid_transforms(n_inst, seed=10)
    ids = [(7 * i + 3) & 0xFFFFFF for i in range(n_inst)]
    masks = [(i % 255) + 1 for i in range(n_inst)]
    flags = [i % 4 for i in range(n_inst)]
    ob = orc.Blas.from_mesh(mesh)
    ot = orc.Tlas([ob] * n_inst, xf, ids=ids, masks=masks, flags=flags)
    gb = ctx.build_blas_from_mesh(mesh)
    gt = ctx.build_tlas([gb] * n_inst, xf, ids=ids, masks=masks, flags=flags, keep_scratch=True)
    ctx.sync()
    np.testing.assert_array_equal(gt.stage("sorted_codes"), ot.sorted_morton())
    np.testing.assert_array_equal(gt.stage("sorted_indices"), ot.perm())
    g, o = T.parse_tlas_blob(gt.blob()), T.parse_tlas_blob(ot.blob())
    np.testing.assert_array_equal(g["header"], o["header"])
    np.testing.assert_array_equal(g["nodes"].view(np.uint8), o["nodes"].view(np.uint8))
    for f in ("w2o", "id_mask", "hg_flags", "o2w", "instance_index"):  # "blas" holds an address: differs by design
        np.testing.assert_array_equal(g["meta"][f], o["meta"][f])
    assert (g["meta"]["blas"] == gb.result.ptr).all()


def test_empty_tlas(ctx, orc):
    """TraceEmptyAccelerationStructure (UT:4000-4008): header + one zero box, every ray misses."""
    gt = ctx.build_tlas([], [])
    ot = orc.Tlas([], [])
    np.testing.assert_array_equal(gt.blob(), ot.blob())
    from helpers import ut_rays
    hits = ctx.trace(gt, ut_rays())
    assert (hits["primitive_index"] == T.NO_HIT).all()


def test_build_argument_errors(ctx, rt):
    import ctypes as C
    info = T.PrebuildInfo()
    mesh = scenes.icosphere(1)
    vb, ib = ctx.upload(mesh.vertices), ctx.upload(mesh.indices)
    d = (T.GeometryDesc * 1)()
    d[0].vertex_buffer, d[0].vertex_count, d[0].vertex_stride_bytes = vb.ptr, mesh.vertices.shape[0], 24
    d[0].index_buffer, d[0].index_count, d[0].index_format = ib.ptr, mesh.indices.size, 32
    rt.check(rt.lib.rt_blas_prebuild(ctx.handle, d, 1, 0, C.byref(info)))
    scratch, result = ctx.alloc(info.scratch_bytes), ctx.alloc(info.result_bytes)
    # result too small -> RT_ERR_TOO_SMALL; null destination -> E_INVALIDARG (FL/GpuBVH2Builder.cpp:145-148)
    assert rt.lib.rt_blas_build(ctx.handle, d, 1, 0, scratch.ptr, scratch.nbytes, result.ptr, 64) == -3
    assert rt.lib.rt_blas_build(ctx.handle, d, 1, 0, scratch.ptr, scratch.nbytes, None, 0) == -1
    d[0].index_format = 8
    assert rt.lib.rt_blas_build(ctx.handle, d, 1, 0, scratch.ptr, scratch.nbytes, result.ptr, result.nbytes) == -1
    assert b"index_format" in rt.lib.rt_last_error()


@pytest.mark.parametrize("n", [255, 256, 257, 511, 512, 513, 767, 1023, 1025, 4097, 65537])
def test_blob_bit_exact_around_fit_block_boundaries(n, ctx, orc):
    """k_fit_local fits 256 sorted slots per thread block in shared memory and hands the boundary-crossing nodes to
    k_fit_exits: sizes at, just below and just above multiples of the block size, for 0 and 1 treelet passes."""
    mesh = scenes.triangle_soup(n, seed=1000 + n, extent=20.0, edge=1.5)
    for flags in (T.BUILD_FLAG_PREFER_FAST_BUILD, 0):
        ref = orc.Blas.from_mesh(mesh, build_flags=flags)
        acc = ctx.build_blas_from_mesh(mesh, build_flags=flags)
        np.testing.assert_array_equal(acc.blob(), ref.blob())


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 513, 4097, 65537, 300001])
def test_fused_fast_build_equals_staged_build(n, ctx):
    """PREFER_FAST_BUILD runs k_lbvh_fit (hierarchy emission + fit in one bottom-up kernel); the same flags with
    ALLOW_UPDATE run the staged chain (k_hierarchy, k_fit_local, k_fit_exits).  Same tree, same bytes: the reference
    blob AND the traversal section (BVH2 wide nodes, packed triangles, 4-wide nodes)."""
    mesh = scenes.triangle_soup(n, seed=4000 + n, extent=30.0, edge=1.0)
    if n == 4097:
        mesh.vertices["position"][: 3 * 2000] = np.tile(mesh.vertices["position"][:3], (2000, 1))  # 2000 equal keys across block boundaries
    fused = ctx.build_blas_from_mesh(mesh, build_flags=T.BUILD_FLAG_PREFER_FAST_BUILD)
    staged = ctx.build_blas_from_mesh(mesh, build_flags=T.BUILD_FLAG_PREFER_FAST_BUILD | T.BUILD_FLAG_ALLOW_UPDATE)
    ctx.sync()
    np.testing.assert_array_equal(fused.blob(), staged.blob())
    f, s = fused.traversal_section(), staged.traversal_section()
    for k in ("wide", "leaf", "wide4"):
        np.testing.assert_array_equal(f[k], s[k], err_msg=k)
    ctx.status()


@pytest.mark.parametrize("kind", ["identical", "two_clusters", "line"])
def test_blob_bit_exact_degenerate_distributions(kind, ctx, orc):
    """Trees the Morton order makes maximally unbalanced or tie-dominated: every triangle identical (all codes equal:
    the hierarchy is decided by the index tie rule alone), two tight clusters far apart, triangles on a line."""
    n = 3000
    mesh = scenes.triangle_soup(n, seed=77, extent=1.0, edge=0.2)
    pos = mesh.vertices["position"]
    if kind == "identical":
        pos[:] = np.tile(pos[:3], (n, 1))
    elif kind == "two_clusters":
        pos[: 3 * (n // 2)] *= np.float32(1e-3)
        pos[3 * (n // 2):] = pos[3 * (n // 2):] * np.float32(1e-3) + np.float32(1000.0)
    else:
        pos[:, 1] = 0.0
        pos[:, 2] = 0.0
    for flags in (T.BUILD_FLAG_PREFER_FAST_BUILD, 0, T.BUILD_FLAG_PREFER_FAST_TRACE):
        ref = orc.Blas.from_mesh(mesh, build_flags=flags)
        acc = ctx.build_blas_from_mesh(mesh, build_flags=flags)
        np.testing.assert_array_equal(acc.blob(), ref.blob())
    # and the traversal section built next to it answers like the oracle's tree
    otlas = orc.Tlas([orc.Blas.from_mesh(mesh)], [scenes.IDENTITY_3X4])
    gtlas = ctx.build_tlas([ctx.build_blas_from_mesh(mesh)], [scenes.IDENTITY_3X4])
    from helpers import random_rays
    lo, hi = pos.min(axis=0) - 1.0, pos.max(axis=0) + 1.0
    rays = random_rays(20000, seed=9, lo=lo, hi=hi)
    ho, hg = otlas.trace(rays, threads=8), ctx.trace(gtlas, rays)
    np.testing.assert_array_equal(hg["t"], ho["t"])


def test_array_of_pointers_layouts(ctx, orc):
    """D3D12_ELEMENTS_LAYOUT_ARRAY_OF_POINTERS (UT:653 multi-geometry BLAS, UT:878-934 TLAS over 50 instances): the same
    bytes as the ARRAY layout, whether the descriptors arrive by value or through pointers."""
    a, b = scenes.icosphere(2), scenes.triangle_soup(700, seed=5, extent=3.0, edge=0.6)
    geoms = []
    for m in (a, b):
        geoms.append(dict(vertices=ctx.upload(m.vertices), vertex_count=m.vertices.shape[0], stride=24, indices=ctx.upload(m.indices),
                          index_count=m.indices.size, index_format=32))
    flat = ctx.build_blas(geoms)
    ptrs = ctx.build_blas(geoms, array_of_pointers=True)
    ref = orc.Blas([dict(vertices=a.vertices, stride=24, indices=a.indices), dict(vertices=b.vertices, stride=24, indices=b.indices)])
    np.testing.assert_array_equal(flat.blob(), ref.blob())
    np.testing.assert_array_equal(ptrs.blob(), ref.blob())
    n_inst = 50
    transforms = scenes.random_rigid_transforms(n_inst, seed=10)
    t_flat = ctx.build_tlas([flat] * n_inst, transforms)
    t_ptrs = ctx.build_tlas([flat] * n_inst, transforms, array_of_pointers=True)
    np.testing.assert_array_equal(t_ptrs.blob(), t_flat.blob())
    from helpers import random_rays
    rays = random_rays(5000, seed=3, lo=(-120, -120, -120), hi=(120, 120, 120))
    np.testing.assert_array_equal(ctx.trace(t_ptrs, rays).view(np.uint8), ctx.trace(t_flat, rays).view(np.uint8))
