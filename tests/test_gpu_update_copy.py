"""GPU parity of the acceleration-structure operations beside the full build, through the C ABI:

* ALLOW_UPDATE / PERFORM_UPDATE (FL/GpuBVH2Builder.cpp:152-204, FL/ComputeAABBs.hlsli:38-67) — restating the reference's
  unit tests UT:1054-1475 and comparing every byte with the oracle's refit;
* CopyRaytracingAccelerationStructure CLONE / COMPACT (FL/GpuBVH2Builder.cpp:330-347) and
  EmitRaytracingAccelerationStructurePostbuildInfo (UT:937-1052);
* the position independence of a bottom-level result buffer (its bytes are the serialised form).
"""
import ctypes as C

import numpy as np
import pytest

from dxrexperiments_b200 import scenes, types as T
from helpers import random_rays

pytestmark = pytest.mark.gpu

UPD = T.BUILD_FLAG_ALLOW_UPDATE


def _mesh_geoms(ctx, mesh):
    vb, ib = ctx.upload(mesh.vertices), ctx.upload(mesh.indices)
    return [dict(vertices=vb, vertex_count=mesh.vertices.shape[0], stride=24, indices=ib, index_count=mesh.indices.size,
                 index_format=32)]


def _wobble(mesh, seed, amp=0.3):
    rng = np.random.Generator(np.random.PCG64(seed))
    out = mesh.vertices.copy()
    out["position"] += rng.uniform(-amp, amp, size=out["position"].shape).astype(np.float32)
    return scenes.Mesh(out, mesh.indices)


def test_updates_allowed_allocate_memory(ctx, rt):
    """UpdatesAllowedAllocateMemoryGpuBVHBuilder (UT:1054-1087): same scratch, result grows by 4n + 4(2n-1)."""
    mesh = scenes.icosphere(2)
    d = rt._geometry_descs(_mesh_geoms(ctx, mesh))
    a, b = T.PrebuildInfo(), T.PrebuildInfo()
    rt.check(rt.lib.rt_blas_prebuild(ctx.handle, d, 1, T.BUILD_FLAG_PREFER_FAST_BUILD, C.byref(a)))
    rt.check(rt.lib.rt_blas_prebuild(ctx.handle, d, 1, T.BUILD_FLAG_PREFER_FAST_BUILD | UPD, C.byref(b)))
    n = mesh.num_triangles
    assert a.scratch_bytes == b.scratch_bytes
    assert b.result_bytes == a.result_bytes + 4 * n + 4 * (2 * n - 1)
    assert a.update_scratch_bytes == 0 and b.update_scratch_bytes > 0
    rt.check(rt.lib.rt_tlas_prebuild(ctx.handle, 50, 0, C.byref(a)))
    rt.check(rt.lib.rt_tlas_prebuild(ctx.handle, 50, UPD, C.byref(b)))
    assert b.result_bytes == a.result_bytes + 4 * 50 + 4 * 99 and a.scratch_bytes == b.scratch_bytes


@pytest.mark.parametrize("name", ["cornell", "icosphere3", "soup5000", "one_triangle", "duplicates"])
def test_allow_update_build_and_caches_bit_exact(name, ctx, orc):
    """StoreSortResultForUpdate / StoreParentIndicesForUpdate (UT:1089-1165) + whole-blob equality with the oracle."""
    dup = scenes.triangle_soup(2000, seed=3, extent=2.0, edge=0.5)
    dup.vertices["position"][: 3 * 300] = np.tile(dup.vertices["position"][:3], (300, 1))
    mesh = {"cornell": scenes.cornell_box(), "icosphere3": scenes.icosphere(3), "soup5000": scenes.triangle_soup(5000, seed=42),
            "one_triangle": scenes.Mesh(scenes.icosphere(0).vertices, scenes.icosphere(0).indices[:3].copy()),
            "duplicates": dup}[name]
    ref = orc.Blas.from_mesh(mesh)
    acc = ctx.build_blas_from_mesh(mesh, build_flags=UPD)
    ctx.status()
    np.testing.assert_array_equal(acc.blob(), ref.blob())                       # ALLOW_UPDATE does not change the blob
    cache, parents = acc.update_caches()
    np.testing.assert_array_equal(cache, ref.sort_cache())
    np.testing.assert_array_equal(parents, ref.parents())
    prims = T.parse_blas_blob(acc.blob())["prims"]
    tri_in = mesh.vertices["position"][mesh.indices.reshape(-1, 3)]
    np.testing.assert_array_equal(prims["v"][cache].reshape(-1, 3, 3), tri_in)  # UT:1114-1123


@pytest.mark.parametrize("subdiv,seed", [(0, 1), (3, 2), (5, 3)])
def test_refit_aabbs_on_update_bit_exact(subdiv, seed, ctx, orc):
    """RefitAABBsOnUpdate (UT:1167-1292): PERFORM_UPDATE with moved vertices == the oracle's refit, byte for byte, and
    traversal of the updated structure agrees with the oracle (t, barycentrics and ids bit-exact)."""
    mesh = scenes.icosphere(subdiv)
    moved = _wobble(mesh, seed)
    ref = orc.Blas.from_mesh(mesh)
    ref.update([dict(vertices=moved.vertices, stride=24, indices=moved.indices)])
    acc = ctx.build_blas_from_mesh(mesh, build_flags=UPD)
    topo0 = T.parse_blas_blob(acc.blob())["nodes"][["flags", "right"]].copy()
    ctx.update_blas(acc, _mesh_geoms(ctx, moved))
    ctx.status()
    np.testing.assert_array_equal(acc.blob(), ref.blob())
    np.testing.assert_array_equal(T.parse_blas_blob(acc.blob())["nodes"][["flags", "right"]], topo0)
    cache, parents = acc.update_caches()                                        # the caches survive an update
    np.testing.assert_array_equal(cache, ref.sort_cache())
    np.testing.assert_array_equal(parents, ref.parents())
    # a second update (back to the original vertices) reproduces the original build exactly
    ot = orc.Tlas([ref], [scenes.IDENTITY_3X4])
    gt = ctx.build_tlas([acc], [scenes.IDENTITY_3X4])
    rays = random_rays(20000, seed, -2.5, 2.5)
    for flags in (0, T.RAY_FLAG_CULL_BACK_FACING_TRIANGLES):
        ho, hg = ot.trace(rays, flags), ctx.trace(gt, rays, flags)
        for f in ("t", "bary", "primitive_index", "instance_index"):
            np.testing.assert_array_equal(hg[f], ho[f])
    ctx.update_blas(acc, _mesh_geoms(ctx, mesh))
    np.testing.assert_array_equal(acc.blob(), orc.Blas.from_mesh(mesh).blob())


def test_update_multi_geometry_with_transform(ctx, orc):
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1]], np.float32)
    idx16 = np.array([0, 1, 2, 2, 1, 3], np.uint16)
    soup = np.ascontiguousarray(scenes.triangle_soup(200, seed=9, extent=5.0).vertices["position"])
    soup2 = (soup + np.float32(0.25)).astype(np.float32)
    xf = np.array([0.5, 0, 0, 1, 0, 2, 0, -3, 0.25, 0, 1, 0.5], np.float32)
    xf2 = np.array([0.5, 0, 0, 2, 0, 2, 0, -1, 0.25, 0, 1, 0.0], np.float32)

    def geoms(soup_pos, x, up):
        o = [dict(vertices=verts, stride=12, indices=idx16, flags=1), dict(vertices=soup_pos, stride=12, indices=None, transform=x, flags=0)]
        g = [dict(vertices=up(verts), vertex_count=4, stride=12, indices=up(idx16), index_count=6, index_format=16, flags=1),
             dict(vertices=up(soup_pos), vertex_count=600, stride=12, indices=None, index_format=0, transform=up(x), flags=0)]
        return o, g

    o0, g0 = geoms(soup, xf, ctx.upload)
    o1, g1 = geoms(soup2, xf2, ctx.upload)
    ref = orc.Blas(o0, build_flags=UPD)
    acc = ctx.build_blas(g0, build_flags=UPD)
    np.testing.assert_array_equal(acc.blob(), ref.blob())
    ref.update(o1)
    ctx.update_blas(acc, g1)
    ctx.status()
    np.testing.assert_array_equal(acc.blob(), ref.blob())


@pytest.mark.parametrize("n_inst", [1, 2, 50, 1000])
def test_tlas_update_bit_exact(n_inst, ctx, orc):
    """SimpleTopLevel..._WithUpdate / TopLevelGpuBVHBuilderWithInstanceTransforms_..._WithUpdate (UT:913-935)."""
    mesh = scenes.icosphere(2)
    xf0 = scenes.random_rigid_transforms(n_inst, seed=10)
    xf1 = scenes.random_rigid_transforms(n_inst, seed=11)
    ids = [(5 * i + 1) & 0xFFFFFF for i in range(n_inst)]
    ob = orc.Blas.from_mesh(mesh)
    ot = orc.Tlas([ob] * n_inst, xf0, build_flags=UPD)
    gb = ctx.build_blas_from_mesh(mesh)
    gt = ctx.build_tlas([gb] * n_inst, xf0, build_flags=UPD)
    cache, parents = gt.update_caches()
    np.testing.assert_array_equal(cache, ot.sort_cache())
    np.testing.assert_array_equal(parents, ot.parents())
    ot.update(xf1, ids=ids)
    ctx.update_tlas(gt, xf1, ids=ids)
    ctx.status()
    g, o = T.parse_tlas_blob(gt.blob()), T.parse_tlas_blob(ot.blob())
    np.testing.assert_array_equal(g["header"], o["header"])
    np.testing.assert_array_equal(g["nodes"].view(np.uint8), o["nodes"].view(np.uint8))
    for f in ("w2o", "id_mask", "hg_flags", "o2w", "instance_index"):
        np.testing.assert_array_equal(g["meta"][f], o["meta"][f])
    rays = random_rays(20000, 7, -60, 60)
    ho, hg = ot.trace(rays), ctx.trace(gt, rays)
    for f in ("t", "bary", "primitive_index", "instance_index", "instance_id"):
        np.testing.assert_array_equal(hg[f], ho[f])
    # the refitted TLAS finds exactly what a fresh build over the moved instances finds
    fresh = ctx.trace(ctx.build_tlas([gb] * n_inst, xf1, ids=ids), rays)
    for f in ("t", "primitive_index", "instance_index"):
        np.testing.assert_array_equal(hg[f], fresh[f])


def test_update_argument_errors(ctx, rt):
    mesh = scenes.icosphere(1)
    g = _mesh_geoms(ctx, mesh)
    plain = ctx.build_blas(g)
    d = rt._geometry_descs(g)
    info = T.PrebuildInfo()
    rt.check(rt.lib.rt_blas_prebuild(ctx.handle, d, 1, UPD, C.byref(info)))
    scratch = ctx.alloc(info.scratch_bytes)
    big = ctx.alloc(info.result_bytes)
    rt.check(rt.lib.rt_as_copy(ctx.handle, big.ptr, big.nbytes, plain.result.ptr, T.COPY_MODE_CLONE))
    # PERFORM_UPDATE without ALLOW_UPDATE; PERFORM_UPDATE on a build that did not allow updates
    assert rt.lib.rt_blas_build(ctx.handle, d, 1, T.BUILD_FLAG_PERFORM_UPDATE, scratch.ptr, scratch.nbytes, big.ptr, big.nbytes) == -1
    assert rt.lib.rt_blas_build(ctx.handle, d, 1, UPD | T.BUILD_FLAG_PERFORM_UPDATE, scratch.ptr, scratch.nbytes, big.ptr, big.nbytes) == -1
    assert b"PERFORM_UPDATE" in rt.lib.rt_last_error()
    # different element count
    upd = ctx.build_blas(g, build_flags=UPD)
    other = _mesh_geoms(ctx, scenes.icosphere(2))
    with pytest.raises(rt.RtError):
        ctx.update_blas(upd, other)
    ctx.status()


# ------------------------------------------------------------------------------------------------ copy / compaction
def test_clone_and_compact_copies(ctx, orc, rt):
    """SimpleTopLevelGpuBVHBuilderWithCopy (UT:889-893, 860-870): the copy holds the same blob and traces identically."""
    mesh = scenes.bunny_scale(3)
    xf = scenes.random_rigid_transforms(6, seed=2)
    gb = ctx.build_blas_from_mesh(mesh, build_flags=UPD)
    gt = ctx.build_tlas([gb] * 6, xf, build_flags=UPD)
    rays = random_rays(20000, 3, -60, 60)
    want = ctx.trace(gt, rays)
    bi = gb.info()
    assert (bi.count, bi.top_level, bi.build_flags) == (mesh.num_triangles, 0, UPD)
    assert bi.total_bytes == gb.result.nbytes and bi.blob_bytes == rt.lib.rt_blob_bytes(mesh.num_triangles, 0)
    assert bi.compacted_bytes == bi.total_bytes - 4 * bi.count - 4 * (2 * bi.count - 1)
    for compact in (False, True):
        b2 = gb.clone(compact=compact)
        np.testing.assert_array_equal(b2.blob(), gb.blob())
        i2 = b2.info()
        assert i2.build_flags == (0 if compact else UPD) and i2.total_bytes == (bi.compacted_bytes if compact else bi.total_bytes)
        t_over_copy = ctx.build_tlas([b2] * 6, xf)                               # a TLAS over the copied BLAS
        t_copy = gt.clone(compact=compact)                                       # a copy of the TLAS itself
        for tl in (t_over_copy, t_copy):
            got = ctx.trace(tl, rays)
            np.testing.assert_array_equal(got.view(np.uint8), want.view(np.uint8))
        np.testing.assert_array_equal(t_copy.blob(), gt.blob())
    # whole-buffer equality for CLONE, including traversal section and caches
    np.testing.assert_array_equal(gb.clone().result.download(np.uint8), gb.result.download(np.uint8))
    # a compacted copy can no longer be updated
    c = gb.clone(compact=True)
    with pytest.raises(rt.RtError):
        ctx.update_blas(c, _mesh_geoms(ctx, mesh))
    ctx.status()


def test_copy_argument_errors(ctx, rt):
    gb = ctx.build_blas_from_mesh(scenes.icosphere(1))
    dst = ctx.alloc(gb.result.nbytes)
    assert rt.lib.rt_as_copy(ctx.handle, dst.ptr, dst.nbytes, gb.result.ptr, 3) == -1      # SERIALIZE: E_INVALIDARG (GpuBVH2Builder.cpp:342-346)
    assert rt.lib.rt_as_copy(ctx.handle, dst.ptr, dst.nbytes, gb.result.ptr, 2) == -1      # VISUALIZATION_DECODE
    assert rt.lib.rt_as_copy(ctx.handle, dst.ptr, 128, gb.result.ptr, 0) == -3
    assert rt.lib.rt_as_copy(ctx.handle, None, 0, gb.result.ptr, 0) == -1
    junk = ctx.alloc(4096).zero()
    assert rt.lib.rt_as_copy(ctx.handle, dst.ptr, dst.nbytes, junk.ptr, 0) == -1            # not an acceleration structure
    info = T.AsInfo()
    assert rt.lib.rt_as_get_info(ctx.handle, junk.ptr, C.byref(info)) == -1


def test_emit_postbuild_info(ctx, rt):
    """EmitRaytracingAccelerationStructurePostBuildInfoTest (UT:937-1052): 70 BLASes of 1..70 triangles; every reported
    size is non-zero, strictly increasing with the triangle count and within the prebuild maximum."""
    soup = scenes.triangle_soup(70, seed=1)
    pos = ctx.upload(np.ascontiguousarray(soup.vertices["position"][soup.indices]))
    accs = [ctx.build_blas([dict(vertices=pos, vertex_count=3 * (i + 1), stride=12, indices=None, index_format=0)]) for i in range(70)]
    sizes = ctx.compacted_sizes(accs)
    assert sizes.shape == (70,)
    for i in range(70):
        assert 0 < sizes[i] <= accs[i].result.nbytes
        assert sizes[i] == accs[i].info().compacted_bytes
        if i:
            assert sizes[i] > sizes[i - 1]
    assert ctx.compacted_sizes([]).size == 0


def test_blas_bytes_are_position_independent(ctx):
    """The serialised form of a BLAS is its result buffer: download, upload somewhere else, build a TLAS over it."""
    mesh = scenes.bunny_scale(3)
    gb = ctx.build_blas_from_mesh(mesh)
    wire = gb.result.download(np.uint8)
    pad = ctx.alloc(4096 + 64 * 7)  # shift the allocation pattern so the new address differs
    moved = ctx.upload(wire)
    assert moved.ptr != gb.result.ptr
    from dxrexperiments_b200 import rtcore
    gb2 = rtcore.Accel(ctx, moved, gb.n, top=False)
    rays = random_rays(20000, 4, -3, 3)
    a = ctx.trace(ctx.build_tlas([gb], [scenes.IDENTITY_3X4]), rays)
    b = ctx.trace(ctx.build_tlas([gb2], [scenes.IDENTITY_3X4]), rays)
    np.testing.assert_array_equal(a.view(np.uint8), b.view(np.uint8))
    del pad
