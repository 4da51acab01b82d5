"""CPU tier: the work schedule of the tensor-core kNN pass (ab_knn_schedule, host-only entry point of the C-ABI).
The enumeration below restates tc_item() of alona_b200/csrc/ab_knn_tc.cu: what CTA b does with (whole rounds, R, S)."""
import ctypes

import pytest


@pytest.fixture(scope='module')
def lib():
    from alona_b200 import build, _lib
    build.build()
    return _lib.load()


def _schedule(lib, blocks, sms, smax):
    out = (ctypes.c_int32 * 3)()
    assert lib.ab_knn_schedule(blocks, sms, smax, out) == 0
    return int(out[0]), int(out[1]), int(out[2])


def _items(full, R, S, sms, n_tiles):
    """(cta, block, first tile, tile count) of every work item, as tc_item() enumerates them"""
    grid = sms if full > 0 else min(sms, max(1, R * S))
    items = []
    for b in range(grid):
        for i in range(full):
            items.append((b, b + i * grid, 0, n_tiles))
        p = b
        while p < R * S:
            j, part = divmod(p, S)
            t0 = part * n_tiles // S
            t1 = (part + 1) * n_tiles // S
            items.append((b, full * grid + j, t0, t1 - t0))
            p += grid
    return grid, items


@pytest.mark.parametrize('blocks', [1, 3, 8, 30, 74, 75, 147, 148, 149, 222, 268, 296, 638, 1276, 1954, 2552, 5103])
@pytest.mark.parametrize('smax', [2, 4])
def test_every_block_tile_pair_is_swept_exactly_once(lib, blocks, smax):
    sms, n_tiles = 148, 97
    full, R, S = _schedule(lib, blocks, sms, min(smax, n_tiles))
    assert 1 <= S <= smax and R >= 0 and full >= 0
    grid, items = _items(full, R, S, sms, n_tiles)
    assert full * grid + R == blocks
    seen = {}
    for _, blk, t0, cnt in items:
        assert 0 <= blk < blocks and cnt >= 1
        for t in range(t0, t0 + cnt):
            seen[(blk, t)] = seen.get((blk, t), 0) + 1
    assert len(seen) == blocks * n_tiles and set(seen.values()) == {1}
    # part lists per tail block fit what ab_knn allocates (4 lists, at most two grids of blocks)
    assert R <= 2 * sms
    # cost in sweeps: the busiest CTA; never worse than whole rounds, within half a sweep of the ideal when it can split
    load = {}
    for b, _, _, cnt in items:
        load[b] = load.get(b, 0) + cnt
    cost = max(load.values()) / n_tiles
    whole = -(-blocks // sms)
    assert cost <= whole + 1e-9
    if blocks >= sms:
        assert cost <= blocks / sms + 0.5 + smax / n_tiles


def test_named_configs_schedule(lib):
    # C3 at 1 / 2 / 4 / 8 GPUs: 5103 / 2552 / 1276 / 638 blocks of 256 queries on 148 SMs
    assert _schedule(lib, 5103, 148, 4) == (34, 71, 2)
    assert _schedule(lib, 2552, 148, 4) == (17, 36, 4)
    assert _schedule(lib, 1276, 148, 4) == (8, 92, 3)
    assert _schedule(lib, 638, 148, 4) == (4, 46, 3)
