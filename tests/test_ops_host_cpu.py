"""Host-side helpers of ops.py that decide launch shapes and buffer sizes (no GPU involved)."""
import pytest

from gaussiangrasper_b200 import ops


@pytest.mark.parametrize("C", list(range(1, 65)) + [65, 70, 71, 72, 100, 128, 129, 200, 256])
def test_channel_blocks_cover_every_channel_in_aligned_pieces(C):
    step = 64
    blocks, same_batch = ops._channel_blocks(C, step)
    assert blocks[0][0] == 0 and blocks[-1][1] == C
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0                                   # contiguous, no overlap
    for c0, c1 in blocks:
        assert 0 < c1 - c0 <= step
        assert c0 % 4 == 0                                # every piece starts on a 16-byte boundary of the row
    if C <= step:
        assert blocks == [(0, C)] and same_batch
    else:
        widths = [c1 - c0 for c0, c1 in blocks]
        assert max(widths) - min(widths) <= 4 + (widths[0] - widths[-1])   # near-equal pieces, remainder in the last
        # hit masks recorded for the first block serve the others only if every block stages the same batch size
        assert same_batch == (all(w > 32 for w in widths) or all(w <= 32 for w in widths))


def test_config4_d64_splits_into_two_wide_blocks_that_share_hit_masks():
    assert ops._channel_blocks(71, 64) == ([(0, 36), (36, 71)], True)


def test_tile_bounds_and_key_bits():
    assert ops.tile_bounds_for(480, 640) == (40, 30, 1)
    assert ops.tile_bounds_for(1, 1) == (1, 1, 1)
    assert ops.tile_bounds_for(1080, 1920) == (120, 68, 1)      # 1080 = 67.5 tiles: the last row is partial
    assert ops.tile_bounds_for(17, 33) == (3, 2, 1)
    assert ops.key_bits_for(1) == 33 and ops.key_bits_for(2) == 33 and ops.key_bits_for(3) == 34
    assert ops.key_bits_for(1200) == 32 + 11 and ops.key_bits_for(8 * 3600) == 32 + 15


def test_capacity_leaves_headroom_and_sh_degree_table():
    for m in (0, 1, 10_000, 2_169_874, 136_433_925):
        cap = ops._capacity_for(m)
        assert cap >= m + 65536 and cap >= int(1.3 * m) and cap < 2 ** 31
    assert [ops.sh_degree_from_bases(b) for b in (1, 4, 9, 16, 25)] == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError, match="SH bases"):
        ops.sh_degree_from_bases(5)
