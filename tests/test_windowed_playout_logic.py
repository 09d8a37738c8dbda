"""CPU check of the LOGIC of the windowed playout (synthesis_b200/csrc/tree.cuh::rollout): a lane-by-lane Python restatement of
the device code against the sequential playout it replaces (RolloutPolicy::eval, synthesis/src/policies/rollout.rs:8-31 with
rand 0.8's gen_range zone rule), on arbitrary word streams.  The GPU parity tests compare the kernel with the oracle on real
ChaCha12 streams, where a rejected word (probability ~n / 2^32) never occurs; here the zone can be lowered artificially so
that rejections, column fills, buffer boundaries and terminal plies all meet inside one window.  No GPU, no oracle."""
import random

import pytest

W, H = 9, 7
ALL = (1 << 63) - 1


def zone_low(n):
    """An artificial zone that depends on n like the real one does: 8-15 % of all words are rejected, differently for every n."""
    return int((0.85 + 0.007 * n) * 2**32)


def zone_true(n):
    return 0xFFFFFFFF - ((0xFFFFFFFF - n + 1) % n)


def won(bb):
    cells = {(c, r) for c in range(W) for r in range(H) if (bb >> (7 * c + r)) & 1}
    for (c, r) in cells:
        for dc, dr in ((1, 0), (0, 1), (1, 1), (1, -1)):
            if all((c + i * dc, r + i * dr) in cells for i in range(4)):
                return True
    return False


def sequential(my, op, words, zone_of):
    """The reference's loop: one word per attempt, the hi-th legal column ascending.  Returns (one-hot index, plies, words used)."""
    pos = k = 0
    while True:
        occ = my | op
        legal = [c for c in range(W) if bin((occ >> (7 * c)) & 0x7F).count("1") < H]
        n = len(legal)
        while True:
            v = words[pos]
            pos += 1
            m = v * n
            if (m & 0xFFFFFFFF) <= zone_of(n):
                break
        col = legal[m >> 32]
        row = bin((occ >> (7 * col)) & 0x7F).count("1")
        mover = my | (1 << (7 * col + row))
        my, op = op, mover
        k += 1
        if won(mover):
            return (2 if k & 1 else 0), k, pos
        if (my | op) == ALL:
            return 1, k, pos


def windowed(my, op, words, zone_of, GL, start_pos=0):
    """tree.cuh::rollout, lane by lane (lists stand for the lanes of a group; ballots are Python sets)."""
    WORDS = 4 * GL
    pos, k = start_pos, 0
    while True:
        occ = my | op
        legal_cols = [c for c in range(W) if bin((occ >> (7 * c)) & 0x7F).count("1") < H]
        n = len(legal_cols)
        zone = zone_of(n)
        in_buf = WORDS - (pos % WORDS)
        w = min(in_buf, GL)
        has = [i < w for i in range(GL)]
        v = [words[pos - start_pos + i] if has[i] else 0 for i in range(GL)]
        m = [v[i] * n for i in range(GL)]
        hi = [x >> 32 for x in m]
        rejected = [has[i] and (m[i] & 0xFFFFFFFF) > zone for i in range(GL)]
        col = [legal_cols[hi[i]] for i in range(GL)]
        below = [sum(1 for j in range(i) if has[j] and col[j] == col[i]) for i in range(GL)]
        row = [bin((occ >> (7 * col[i])) & 0x7F).count("1") + below[i] for i in range(GL)]
        overflow = [has[i] and row[i] >= 7 for i in range(GL)]
        fills = [has[i] and row[i] == 6 for i in range(GL)]
        bad = [i for i in range(GL) if rejected[i] or overflow[i] or not has[i]]
        first_bad = bad[0] if bad else GL
        fill_lanes = [i for i in range(GL) if fills[i] and i < first_bad]
        end = first_bad
        filled_last = False
        if fill_lanes and fill_lanes[0] + 1 <= end:
            end = fill_lanes[0] + 1
            filled_last = True
        bit = [(1 << (7 * col[i] + row[i])) if (has[i] and row[i] < 7) else 0 for i in range(GL)]
        x = list(bit)
        o = 2
        while o < GL:  # inclusive OR-scan over the lanes of the same parity
            prev = list(x)
            for i in range(GL):
                if i >= o:
                    x[i] |= prev[i - o]
            o <<= 1
        x = [x[i] | (op if i & 1 else my) for i in range(GL)]
        win = [i < end and won(x[i]) for i in range(GL)]
        full = [i < end and bin(occ).count("1") + i + 1 == 63 for i in range(GL)]
        term = [i for i in range(GL) if win[i] or full[i]]
        if term:
            tl = term[0]
            k += tl + 1
            pos += tl + 1
            return ((2 if k & 1 else 0) if win[tl] else 1), k, pos - start_pos
        if end > 0:
            last = x[end - 1]
            prev_b = x[end - 2] if end >= 2 else op
            my, op = prev_b, last
        k += end
        pos += end + (1 if (not filled_last and end == first_bad and end < w and rejected[end]) else 0)


def random_position(rng, plies):
    my = op = 0
    for _ in range(plies):
        occ = my | op
        legal = [c for c in range(W) if bin((occ >> (7 * c)) & 0x7F).count("1") < H]
        if not legal:
            break
        c = rng.choice(legal)
        r = bin((occ >> (7 * c)) & 0x7F).count("1")
        mover = my | (1 << (7 * c + r))
        if won(mover):
            continue  # keep the position non-terminal: try another move next round
        my, op = op, mover
        if (my | op) == ALL:
            return None
    return my, op


@pytest.mark.parametrize("GL", [16, 32])
@pytest.mark.parametrize("zone_kind", ["true", "low"])
def test_windowed_playout_equals_the_sequential_one(GL, zone_kind):
    rng = random.Random(1234 + GL + (7 if zone_kind == "low" else 0))
    # "low": a tenth of all words are rejected, so rejections land inside windows, next to fills and buffer boundaries
    zone_of = zone_true if zone_kind == "true" else zone_low
    checked = windows_crossing_boundary = 0
    for trial in range(400):
        pos0 = random_position(rng, rng.randrange(0, 50))
        if pos0 is None:
            continue
        my, op = pos0
        words = [rng.getrandbits(32) for _ in range(400)]
        want = sequential(my, op, words, zone_of)
        start = rng.randrange(0, 4 * GL)  # where in the ring the stream stands: windows get cut at the ring's end
        got = windowed(my, op, words, zone_of, GL, start_pos=start)
        assert got == want, (trial, GL, zone_kind, start, hex(my), hex(op), want, got)
        checked += 1
        windows_crossing_boundary += (start % (4 * GL)) + want[2] > 4 * GL
    assert checked > 300 and windows_crossing_boundary > 5


def test_true_zone_matches_rand_0_8():
    """UniformInt::<u8>::sample_single widens to u32: zone = u32::MAX - (u32::MAX - range + 1) % range (rand 0.8.5 uniform.rs)."""
    assert [zone_true(n) for n in range(1, 10)] == [0xFFFFFFFF - ((2**32 - n) % n) for n in range(1, 10)]
    assert zone_true(9) == 2**32 - 5 and zone_true(8) == 2**32 - 1 and zone_true(7) == 2**32 - 5
