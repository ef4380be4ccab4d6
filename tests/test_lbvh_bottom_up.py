"""CPU check of the claim k_lbvh_fit (csrc/build.cu) rests on: the reference's Karras hierarchy
(FL/BuildBVHSplits.hlsli:35-143, restated in oracle_build.cpp `build_hierarchy`) is the binary radix tree of the sorted keys, and
that tree — WITH Karras' node numbering — can be grown from the leaves without any search:

  a finished node covering the sorted slots [lo, hi] is the LEFT child of the node that splits at hi when
  delta(hi, hi+1) > delta(lo-1, lo), otherwise the RIGHT child of the node that splits at lo-1 (delta = -1 outside [0, n));
  a left child is internal node `hi`, a right child internal node `lo`, the range [0, n-1] is node 0.

The walk below is a plain-Python restatement of that rule (no GPU, no product code); its {parent, left, right} arrays must equal
the oracle's hierarchy bit for bit, also with duplicate Morton codes (ties decided by `clz(i ^ j) + 31`) and at the block-boundary
sizes of the CUDA kernel.  The GPU side is tests/test_gpu_build.py::test_fused_fast_build_equals_staged_build.
"""
import numpy as np
import pytest

from dxrexperiments_b200 import scenes, types as T


def _clz32(x: int) -> int:
    return 32 - int(x).bit_length()


def _delta(codes, n, i):
    """delta(i, i + 1) of the reference (BuildBVHSplits.hlsli:48-64); -1 outside [0, n - 1)."""
    if i < 0 or i >= n - 1:
        return -1
    a, b = int(codes[i]), int(codes[i + 1])
    return _clz32(a ^ b) if a != b else _clz32(i ^ (i + 1)) + 31


def bottom_up_hierarchy(codes: np.ndarray):
    n = len(codes)
    n_int = n - 1
    parent = np.zeros(2 * n - 1, np.uint32)
    left = np.zeros(2 * n - 1, np.uint32)
    right = np.zeros(2 * n - 1, np.uint32)
    waiting = {}  # split -> (side, node, lo, hi) of the child that arrived first
    work = [(n_int + s, s, s) for s in range(n)]  # (node, lo, hi): every leaf is finished
    while work:
        node, lo, hi = work.pop()
        if lo == 0 and hi == n - 1:
            assert node == 0 or n == 1
            continue
        go_right = _delta(codes, n, hi) > _delta(codes, n, lo - 1)
        assert _delta(codes, n, hi) != _delta(codes, n, lo - 1)  # never equal for distinct (code, index) keys
        split = hi if go_right else lo - 1
        if split not in waiting:
            waiting[split] = (go_right, node, lo, hi)
            continue
        o_right, o_node, o_lo, o_hi = waiting.pop(split)
        assert o_right != go_right
        l_node, r_node = (node, o_node) if go_right else (o_node, node)
        nlo, nhi = (lo, o_hi) if go_right else (o_lo, hi)
        if nlo == 0 and nhi == n - 1:
            pid = 0
        else:
            pid = nhi if _delta(codes, n, nhi) > _delta(codes, n, nlo - 1) else nlo
        left[pid], right[pid] = l_node, r_node
        parent[l_node] = parent[r_node] = pid
        work.append((pid, nlo, nhi))
    assert not waiting
    return parent, left, right


CASES = {
    "soup_300": lambda: scenes.triangle_soup(300, seed=11, extent=10.0, edge=1.0),
    "soup_1025": lambda: scenes.triangle_soup(1025, seed=12, extent=40.0, edge=1.0),   # crosses four 256-slot blocks
    "soup_257": lambda: scenes.triangle_soup(257, seed=13, extent=5.0, edge=0.5),
    "icosphere3": lambda: scenes.icosphere(3),
    "two": lambda: scenes.Mesh(scenes.icosphere(0).vertices, scenes.icosphere(0).indices[:6].copy()),
}


def _duplicates():
    m = scenes.triangle_soup(700, seed=3, extent=2.0, edge=0.5)
    m.vertices["position"][: 3 * 300] = np.tile(m.vertices["position"][:3], (300, 1))  # 300 identical keys: the index tie rule alone
    return m


CASES["duplicates"] = _duplicates


@pytest.mark.parametrize("name", sorted(CASES))
def test_bottom_up_rule_reproduces_the_karras_hierarchy(name, orc):
    mesh = CASES[name]()
    ref = orc.Blas.from_mesh(mesh, build_flags=T.BUILD_FLAG_PREFER_FAST_BUILD)  # no treelet pass: the plain Karras tree
    codes = ref.sorted_morton()
    n = ref.n
    assert (np.diff(codes.astype(np.int64)) >= 0).all()
    parent, left, right = bottom_up_hierarchy(codes)
    h = ref.hierarchy()
    np.testing.assert_array_equal(left[: n - 1], h["left"][: n - 1])
    np.testing.assert_array_equal(right[: n - 1], h["right"][: n - 1])
    np.testing.assert_array_equal(parent[1:], h["parent"][1:] & 0x7FFFFFFF)


def test_bottom_up_rule_on_adversarial_keys():
    """Hand-made key sequences: all equal, strictly doubling, one outlier — compared with a direct restatement of Karras' search."""
    def karras(codes):
        n = len(codes)

        def lcp(a, b):
            if a < 0 or b < 0 or a >= n or b >= n:
                return -1
            ca, cb = int(codes[a]), int(codes[b])
            return _clz32(ca ^ cb) if ca != cb else _clz32(a ^ b) + 31
        left = np.zeros(n - 1, np.uint32)
        right = np.zeros(n - 1, np.uint32)
        for i in range(n - 1):
            d = 1 if lcp(i, i + 1) - lcp(i, i - 1) > 0 else -1
            mn = lcp(i, i - d)
            ln = 0
            while lcp(i, i + (ln + 1) * d) > mn:
                ln += 1
            j = i + ln * d
            first, last = min(i, j), max(i, j)
            cp = lcp(first, last)
            split = first
            while split + 1 < last and lcp(first, split + 1) > cp:
                split += 1
            left[i] = (n - 1 + split) if split == first else split
            right[i] = (n - 1 + split + 1) if split + 1 == last else split + 1
        return left, right
    for codes in (np.zeros(37, np.uint32), (1 << np.arange(30)).astype(np.uint32), np.array([5] * 9 + [1 << 29] + [(1 << 29) + 1] * 6, np.uint32),
                  np.sort(np.random.Generator(np.random.PCG64(1)).integers(0, 1 << 30, 513, dtype=np.uint32) & np.uint32(0x3FFFFF00))):
        _, left, right = bottom_up_hierarchy(codes)
        kl, kr = karras(codes)
        n = len(codes)
        np.testing.assert_array_equal(left[: n - 1], kl)
        np.testing.assert_array_equal(right[: n - 1], kr)
