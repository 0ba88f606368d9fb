"""Bit tricks and work distribution of the two-pass ingest (arroyo_b200/csrc/ingest_two_pass.cuh, window_agg.cu),
restated in Python and checked exhaustively / on random inputs without a GPU:

  * agg_kernel's accumulators: the 64-bit wrapping SUM kept as a 32-bit low word plus ONE word that packs the row count
    (low 16 bits) with the signed sum of carries minus negative rows (high 16 bits), a third word only for values that
    are not sign-extended 32-bit numbers -- must reproduce (count, sum mod 2^64) for any values, up to the number of
    rows a block aggregates between two flushes;
  * the lookup's zero-byte trick on eight 8-bit tags: may flag false positives, never misses a match; with the exact
    re-check the candidates are exactly the matching slots;
  * the slot decode of the flagged bits;
  * the work items of agg_kernel (whole buckets round-robin, the last ragged round cut into part-buckets): every row of
    every bucket is covered exactly once and no block gets two items of the last round.
(The GPU tests compare the kernels themselves with the oracle: tests/test_gpu_two_pass.py, test_gpu_fullsize.py.)"""
import random

M32, M64 = (1 << 32) - 1, (1 << 64) - 1
FLUSH_ROWS = 31 * 16 * 64  # P2_FLUSH_ITERS x P2_NW x P2_CH


def packed_accumulate(vals):
    lo = hi = cw = 0
    for v in vals:
        u = v & M64
        vl, vh = u & M32, u >> 32
        narrow = (M32 if vl >> 31 else 0) == vh
        old, lo = lo, (lo + vl) & M32
        d = 1 if ((old + vl) & M32) < vl else 0
        if narrow:
            d = (d - (vl >> 31)) & M32
        else:
            hi = (hi + vh) & M32
        cw = (cw + 1 + ((d << 16) & M32)) & M32
    c = cw & 0xFFFF
    x = (cw - c) & M32
    x -= (1 << 32) if x >> 31 else 0
    return c, ((((hi + ((x >> 16) & M32)) & M32) << 32) + lo)


def test_packed_row_count_and_carry_word_reproduces_count_and_wrapping_sum():
    rng = random.Random(1)
    shapes = {
        "small": lambda: rng.randint(-2**31, 2**31 - 1),
        "wide": lambda: rng.randint(-2**63, 2**63 - 1),
        "edges": lambda: rng.choice([-1, -2**31, 2**31 - 1, 0, 2**32 - 1, -2**32, 2**31, -2**31 - 1, 2**63 - 1, -2**63]),
        "negative": lambda: -rng.randint(0, 2**31),
        "all_carry": lambda: 2**32 - 1,
    }
    for name, draw in shapes.items():
        for n in (1, 2, 300, FLUSH_ROWS):
            vals = [draw() for _ in range(n)]
            c, s = packed_accumulate(vals)
            assert (c, s) == (n, sum(vals) & M64), (name, n)


def zero_byte_candidates(tags, tag):
    """m of agg_kernel's lookup: bit 8j = slot j (0..3), bit 8j + 4 = slot 4 + j."""
    tag4 = tag * 0x01010101
    w0 = int.from_bytes(bytes(tags[0:4]), "little") ^ tag4
    w1 = int.from_bytes(bytes(tags[4:8]), "little") ^ tag4
    z = lambda x: ((x - 0x01010101) & ~x & 0x80808080) & M32  # noqa: E731
    return (z(w0) >> 7) | (z(w1) >> 3)


def decode(bit):
    return (bit >> 3) + (bit & 4)


def test_zero_byte_trick_never_misses_a_tag_and_the_exact_recheck_removes_every_false_positive():
    rng = random.Random(2)
    false_positives = 0
    for _ in range(20000):
        tag = rng.randint(1, 255)
        tags = [rng.choice([0, tag, tag ^ 1, rng.randint(0, 255)]) for _ in range(8)]
        m = zero_byte_candidates(tags, tag)
        flagged = set()
        while m:
            bit = (m & -m).bit_length() - 1
            m &= m - 1
            flagged.add(decode(bit))
        true = {j for j in range(8) if tags[j] == tag}
        assert true <= flagged                                   # no match is missed
        assert {j for j in flagged if tags[j] == tag} == true    # the kernel's exact re-check
        false_positives += len(flagged - true)
    assert false_positives > 0  # they do occur (a byte above a matching one that differs in bit 0): hence the re-check


def test_slot_decode_covers_the_eight_slots():
    assert sorted(decode(b) for b in (0, 8, 16, 24, 4, 12, 20, 28)) == list(range(8))


def agg_work_items(n_buckets, max_blocks, rows_per_bucket, slices=1):
    """launch_two_pass's tail_first / tail_slices and agg_kernel's mapping work -> (bucket, slice, n_slices)."""
    tail_first, tail_slices = n_buckets, 1
    if slices == 1 and n_buckets > max_blocks and rows_per_bucket >= 2048:
        rest = n_buckets % max_blocks
        if rest and max_blocks // rest >= 2:
            tail_first, tail_slices = n_buckets - rest, min(max_blocks // rest, 4)
    head = tail_first * slices
    n_work = head + (n_buckets - tail_first) * tail_slices
    items = []
    for work in range(n_work):
        if work >= head:
            items.append((tail_first + (work - head) // tail_slices, (work - head) % tail_slices, tail_slices))
        else:
            items.append((work // slices, work % slices, slices))
    return items, min(n_work, max_blocks)


def test_work_items_cover_every_chunk_of_every_bucket_once_and_the_last_round_is_not_ragged():
    for n_buckets, max_blocks, rows in ((1024, 296, 16384), (1024, 296, 100), (300, 296, 50000), (64, 296, 262144),
                                        (592, 296, 20000), (1000, 296, 16000), (1, 296, 10)):
        slices = 1 if n_buckets >= max_blocks else max(1, min((max_blocks + n_buckets - 1) // n_buckets, rows // 4096))
        items, grid = agg_work_items(n_buckets, max_blocks, rows, slices)
        covered = {}
        for b, s, ns in items:
            chunks = (rows + 63) // 64
            lo, hi = chunks * s // ns, chunks * (s + 1) // ns
            covered.setdefault(b, []).append((lo, hi))
        assert sorted(covered) == list(range(n_buckets))
        for b, spans in covered.items():
            spans.sort()
            assert spans[0][0] == 0 and spans[-1][1] == (rows + 63) // 64
            assert all(a[1] == c[0] for a, c in zip(spans, spans[1:]))
        # blocks take items round-robin: the last round holds at most one item per block
        rounds = (len(items) + grid - 1) // grid
        assert len(items) - (rounds - 1) * grid <= grid
    # the headline: 1024 buckets over 296 blocks = three full rounds + 136 buckets cut in two = 272 half-items
    items, grid = agg_work_items(1024, 296, 16384)
    assert len(items) == 888 + 272 and grid == 296
