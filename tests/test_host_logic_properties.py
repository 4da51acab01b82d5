"""CPU tier: property tests (hypothesis) of the host-side logic the sharded path relies on."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from alona_b200.pipeline import pack_rows, prune_threshold, reshard_plan, shard_bounds, unpack_rows


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 100), st.floats(0.0, 0.999, allow_nan=False))
def test_prune_threshold_is_the_smallest_kept_overlap(k, prune):
    """alona/clustering.py:225-227 keeps an edge iff v/(k+(k-v)) > prune_snn; the kernel keeps overlap >= v_min."""
    v_min = prune_threshold(k, prune)
    for v in range(0, k + 1):
        assert (v / (k + (k - v)) > prune) == (v >= v_min)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 5_000_000), st.integers(1, 16))
def test_shard_bounds_partition_the_rows(n, world):
    b = shard_bounds(n, world)
    assert b[0] == 0 and b[-1] == n and len(b) == world + 1
    sizes = np.diff(b)
    assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 40), min_size=1, max_size=6), st.integers(1, 4))
def test_pack_unpack_roundtrip_for_ragged_shards(counts, width):
    parts = [torch.arange(c * width, dtype=torch.float64).reshape(c, width) + 1000 * r for r, c in enumerate(counts)]
    n_max = max(counts)
    gathered = torch.cat([pack_rows(p, n_max) for p in parts], dim=0)
    assert torch.equal(unpack_rows(gathered, counts), torch.cat(parts, dim=0))


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 500), min_size=1, max_size=8))
def test_reshard_plan_moves_every_row_exactly_once_in_order(counts):
    world = len(counts)
    plan, bounds = reshard_plan(counts, world)
    assert bounds == shard_bounds(sum(counts), world)
    first = np.concatenate([[0], np.cumsum(counts)])
    for t in range(world):
        got = []
        for r in range(world):                                    # rows arrive in source-rank order
            s, c = plan[r][t]
            got.extend(range(int(first[r]) + s, int(first[r]) + s + c))
        assert got == list(range(bounds[t], bounds[t + 1]))
