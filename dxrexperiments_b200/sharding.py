"""Multi-GPU partitioning of a progressive frame.

The frame shards by SAMPLE INDEX and/or by SCREEN TILE; every rank holds a replicated BVH.  The shaders'
random numbers are a pure function of (pixel, frameCount), so rank r rendering global samples
``r, r + world, ...`` with ``frameCount = global sample index`` reproduces exactly the samples a single GPU
would have rendered; only the order of the fp32 summation differs.

Each rank keeps the reference's running mean over ITS samples (RayGen: ProgressiveRaytracing.hlsl:36-38).
To combine, one weighted sum-reduce adds the buffers onto the root: ``rt_accum_reduce`` (NCCL, the weight
``n_rank / n_total`` applied inside the reduction) on GPUs, gloo in the CPU tests.

:func:`plan` combines both axes: ``world = strip_groups x sample_groups``.  Rank ``r`` belongs to strip group
``r % strip_groups`` (it renders every ``strip_groups``-th horizontal strip, ``rt_dispatch_rays_interleaved``, a balanced
split of sky and geometry) and to sample group ``r // strip_groups`` (it renders that group's round-robin share of the
samples).  Its buffer is zero outside its strips, so the same reduce assembles the frame.
"""
from dataclasses import dataclass
from typing import List, Tuple


@dataclass
class ShardPlan:
    rank: int
    world: int
    strip_groups: int
    strip_group: int
    sample_groups: int
    sample_group: int
    samples: List[int]   # global sample indices (= frameCount values) this rank renders
    weight: float        # this rank's weight in the reduce: len(samples) / total_samples
    strip_rows: int

    def rows(self, height: int) -> List[int]:
        """Image rows this rank renders."""
        return [y for y in range(height) if (y // self.strip_rows) % self.strip_groups == self.strip_group]


def default_strip_groups(world: int, total_samples: int) -> int:
    """Shard by sample index as far as the samples go (identical work on every rank), by strips beyond that."""
    g = 1
    while world // g > max(total_samples, 1) or world % g:
        g += 1
    return g


def plan(rank: int, world: int, total_samples: int, strip_groups: int = None, strip_rows: int = 32) -> ShardPlan:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if total_samples <= 0:
        raise ValueError("total_samples must be positive")
    if strip_groups is None:
        strip_groups = default_strip_groups(world, total_samples)
    if strip_groups < 1 or world % strip_groups:
        raise ValueError("world must be a multiple of strip_groups")
    if strip_rows < 4 or strip_rows & (strip_rows - 1):
        raise ValueError("strip_rows must be a power of two >= 4")
    sample_groups = world // strip_groups
    sg, tg = rank // strip_groups, rank % strip_groups
    samples = list(range(sg, total_samples, sample_groups))
    return ShardPlan(rank, world, strip_groups, tg, sample_groups, sg, samples, len(samples) / float(total_samples), strip_rows)


def samples_for_rank(rank: int, world: int, total_samples: int) -> List[int]:
    """Global sample indices rendered by `rank` (round-robin, so every rank's set is contiguous in time)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, total_samples, world))


def combine_scale(rank: int, world: int, total_samples: int) -> float:
    """Factor that turns rank's running mean into its share of the global mean."""
    if total_samples <= 0:
        raise ValueError("total_samples must be positive")
    return len(samples_for_rank(rank, world, total_samples)) / float(total_samples)


def tiles_for_rank(rank: int, world: int, width: int, height: int, tile: int = 64) -> List[Tuple[int, int, int, int]]:
    """Screen-tile sharding: (x0, y0, x1, y1) rectangles of `rank`, tiles dealt round-robin in raster order.
    Pixels outside a rank's tiles stay zero in its buffer, so the same sum-reduce assembles the frame."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    out, k = [], 0
    for y0 in range(0, height, tile):
        for x0 in range(0, width, tile):
            if k % world == rank:
                out.append((x0, y0, min(x0 + tile, width), min(y0 + tile, height)))
            k += 1
    return out


@dataclass
class BandPlan:
    """Realtime frame (RealtimeRaytracingPipeline + DenoiseCompositor) sharded by contiguous row bands."""
    rank: int
    world: int
    y0: int   # core rows [y0, y1): what this rank contributes to the composited frame
    y1: int
    r0: int   # rows [r0, r1) this rank renders and filters: the core plus the filter's reach on both sides
    r1: int


def band_plan(rank: int, world: int, height: int, halo: int) -> BandPlan:
    """Rank `rank` of `world` owns rows [y0, y1) of a `height`-row frame and renders / denoises `halo` more rows on each
    side (clipped at the image border): the separable bilateral filter reaches maxKernelSize rows up and down
    (BilateralFilter.hlsli:92-115), so with halo >= maxKernelSize the core rows of a band's filter output are the rows
    the full-frame filter produces, bit for bit.  The core bands partition the frame; a rank's buffer is zero outside
    its core, so the weight-1 sum of rt_accum_reduce composites the frame."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if height <= 0 or halo < 0:
        raise ValueError("height must be positive and halo non-negative")
    y0, y1 = height * rank // world, height * (rank + 1) // world
    return BandPlan(rank, world, y0, y1, max(0, y0 - halo), min(height, y1 + halo))
