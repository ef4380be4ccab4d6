"""Multi-GPU partitioning of a progressive frame.

The frame shards by SAMPLE INDEX and/or by SCREEN TILE; every rank holds a replicated BVH.  The shaders'
random numbers are a pure function of (pixel, frameCount), so rank r rendering global samples
``r, r + world, ...`` with ``frameCount = global sample index`` reproduces exactly the samples a single GPU
would have rendered; only the order of the fp32 summation differs.

Each rank keeps the reference's running mean over ITS samples (RayGen: ProgressiveRaytracing.hlsl:36-38).
To combine, a rank scales its buffer by ``n_rank / n_total`` (``rt_scale_buffer``) and one sum-reduce
(NCCL on GPUs, gloo in the CPU tests) adds the buffers onto the root.
"""
from typing import List, Tuple


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
