"""CPU model of the register z scan (csrc/edt_scan_registers.cuh), statement by statement for the
parts that changed when lines of 1025 .. 2048 voxels got two words per lane: the per-word tables
of every half by warp scans (inclusive prefix maximum / suffix minimum over 32 lanes), the totals
that cross from one half to the other, and the per-voxel searches (one to the left of a lane's
first voxel, one to the right of its last, +1 / restart inside the lane's four voxels). Checked
against the definition: the squared distance to the nearest voxel of the opposite class along
the line, NONE when the line has a single class (sdfgen.hpp:57-74 + the z loop of
sdfgen.cpp:354-390 for both fields at once)."""
import random

FAR = 1 << 20
FAR_THRESHOLD = 1 << 19
NONE = 0x7FFFFFFF
WARP = 32


def valid_bits(word, length):
    remaining = length - (word << 5)
    return 0xFFFFFFFF if remaining >= 32 else ((1 << remaining) - 1 if remaining > 0 else 0)


def prefix_max(values):
    out = list(values)
    offset = 1
    while offset < WARP:
        out = [max(out[lane], out[lane - offset]) if lane >= offset else out[lane]
               for lane in range(WARP)]
        offset <<= 1
    return out


def suffix_min(values):
    out = list(values)
    offset = 1
    while offset < WARP:
        out = [min(out[lane], out[lane + offset]) if lane + offset < WARP else out[lane]
               for lane in range(WARP)]
        offset <<= 1
    return out


def scan_line(classes):
    """classes: list of 0 / 1 (1 = filled), len a multiple of 4 and <= 2048.
    Returns the packed words the kernel stores: class << 31 | squared distance."""
    length = len(classes)
    assert length % 4 == 0 and 0 < length <= 2048
    iterations = 1
    while iterations * 128 < length:
        iterations *= 2
    num_words = (length + 31) >> 5
    halves = (iterations + 7) // 8
    words = [0] * (iterations * 4)
    for index, value in enumerate(classes):
        words[index >> 5] |= value << (index & 31)
    # step 2: lane w of half h owns word 32 h + w
    last = {"filled": [], "free": []}
    first = {"filled": [], "free": []}
    for h in range(halves):
        lf, lr, ff, fr = [], [], [], []
        for lane in range(WARP):
            w = (h << 5) + lane
            mine = words[w] if w < len(words) else 0
            valid = valid_bits(w, length) if w < num_words else 0
            filled_bits, free_bits = mine & valid, ~mine & valid & 0xFFFFFFFF
            lf.append((w << 5) + filled_bits.bit_length() - 1 if filled_bits else -FAR)
            lr.append((w << 5) + free_bits.bit_length() - 1 if free_bits else -FAR)
            ff.append((w << 5) + (filled_bits & -filled_bits).bit_length() - 1 if filled_bits else FAR)
            fr.append((w << 5) + (free_bits & -free_bits).bit_length() - 1 if free_bits else FAR)
        last["filled"].append(prefix_max(lf))
        last["free"].append(prefix_max(lr))
        first["filled"].append(suffix_min(ff))
        first["free"].append(suffix_min(fr))
    before = {"filled": [], "free": []}
    after = {"filled": [], "free": []}
    preceding = {"filled": -FAR, "free": -FAR}
    for h in range(halves):
        for kind in ("filled", "free"):
            following = first[kind][h + 1][0] if h + 1 < halves else FAR
            merged_last = [max(v, preceding[kind]) for v in last[kind][h]]
            merged_first = [min(v, following) for v in first[kind][h]]
            before[kind].append([preceding[kind]] + merged_last[:-1])      # shuffle up by one
            after[kind].append(merged_first[1:] + [following])             # shuffle down by one
            preceding[kind] = max(preceding[kind], last[kind][h][WARP - 1])
    # step 3
    out = [0] * length
    for it in range(iterations):
        for lane in range(WARP):
            vector_index = (it << 5) + lane
            if vector_index >= length >> 2:
                continue
            group, bit0 = lane >> 3, (lane & 7) << 2
            w = (it << 2) + group
            word_start = w << 5
            half, source = it >> 3, w & 31
            left_of_filled = before["free"][half][source] - word_start
            left_of_free = before["filled"][half][source] - word_start
            right_of_filled = after["free"][half][source] - word_start
            right_of_free = after["filled"][half][source] - word_start
            word = words[w]
            valid_here = valid_bits(w, length)
            opposite_of_filled = ~word & valid_here & 0xFFFFFFFF
            opposite_of_free = word & valid_here
            nibble = (word >> bit0) & 0xF
            flips = nibble ^ (nibble >> 1)
            first_is_filled, last_is_filled = bool(nibble & 1), bool(nibble & 8)
            below = (opposite_of_filled if first_is_filled else opposite_of_free) & ((1 << bit0) - 1)
            above = ((opposite_of_filled if last_is_filled else opposite_of_free) >> (bit0 + 3)) >> 1
            left_outside = left_of_filled if first_is_filled else left_of_free
            right_outside = right_of_filled if last_is_filled else right_of_free
            left = [0] * 4
            right = [0] * 4
            left[0] = bit0 - (below.bit_length() - 1 if below else left_outside)
            right[3] = ((bit0 + 3 + (above & -above).bit_length()) if above else right_outside) - (bit0 + 3)
            for k in range(1, 4):
                left[k] = 1 if (flips >> (k - 1)) & 1 else left[k - 1] + 1
            for k in range(2, -1, -1):
                right[k] = 1 if (flips >> k) & 1 else right[k + 1] + 1
            for k in range(4):
                nearest = min(left[k], right[k])
                squared = NONE if nearest >= FAR_THRESHOLD else nearest * nearest
                out[4 * vector_index + k] = (((nibble >> k) & 1) << 31) | squared
    return out


def brute_line(classes):
    length = len(classes)
    out = []
    for q in range(length):
        best = None
        for i in range(length):
            if classes[i] != classes[q] and (best is None or abs(i - q) < best):
                best = abs(i - q)
        out.append((classes[q] << 31) | (NONE if best is None else best * best))
    return out


def test_scan_model_matches_the_definition():
    rng = random.Random(5)
    lengths = [4, 32, 36, 128, 132, 260, 512, 1000, 1024, 1028, 1100, 1536, 2044, 2048]
    for length in lengths:
        cases = [[0] * length, [1] * length]
        for position in (0, 1023, 1024, 1055, 1056, length - 1):
            if position < length:
                line = [0] * length
                line[position] = 1
                cases.append(line)
                cases.append([1 - v for v in line])
        if length > 1024:
            cases.append([1] * 1024 + [0] * (length - 1024))
        for density in (0.002, 0.02, 0.5):
            cases.append([1 if rng.random() < density else 0 for _ in range(length)])
        for line in cases:
            want = brute_line(line) if length <= 300 else None
            got = scan_line(line)
            if want is None:
                # the definition by two sweeps (exact, O(n)) for the long lines
                nearest = [None] * length
                for sweep in (range(length), range(length - 1, -1, -1)):
                    seen = {0: None, 1: None}
                    for q in sweep:
                        seen[line[q]] = q
                        other = seen[1 - line[q]]
                        if other is not None:
                            d = abs(q - other)
                            nearest[q] = d if nearest[q] is None else min(nearest[q], d)
                want = [(line[q] << 31) | (NONE if nearest[q] is None else nearest[q] ** 2)
                        for q in range(length)]
            assert got == want, (length, line[:64])
