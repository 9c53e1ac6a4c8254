"""CPU: pure-Python models of the line algorithm the CUDA envelope kernels run
(voxelized_geometry_tools_b200/csrc/edt_envelope_inplace.cuh and edt_envelope_lean.cuh), checked
against brute force.

The kernel cannot run without a GPU, but its algorithm can: this model mirrors the kernel's two
phases statement by statement (same stack discipline, same integer pop test, same class-bit run
tracking), so a logic error shows up here on CPU before any GPU time is spent.

Line semantics: each voxel has a class bit and a partial squared distance to the opposite class
(NONE = none yet). The transform of the line gives, for voxel q,
    min( min over same-run voxels v with a value of (q-v)^2 + value(v),
         (q-(a-1))^2 if the run starts at a > 0, ((b+1)-q)^2 if it ends at b < n-1 ).
"""
import random

NONE = 0x7FFFFFFF
NO_HEIGHT = 0x3FFFFFFF


def hidden(below, top, incoming):
    # MiddleIsHidden: crossing(below, top) >= crossing(top, incoming), cross-multiplied.
    return (top[1] - below[1]) * (incoming[0] - top[0]) >= (incoming[1] - top[1]) * (top[0] - below[0])


def kernel_line(classes, values):
    n = len(classes)
    num_words = (n + 31) // 32
    rows = [None] * n                 # the line's own storage, reused as the stack
    class_words = [0] * num_words
    # ---- phase 1: per run, stack = [implicit zero site left of the run] + stored own sites;
    #      the zero site right of the run pops what it hides when the run ends.
    slot = depth = 0
    has_left = False
    left_zero = top = below = (0, 0)
    previous = 0

    def pop_hidden(incoming):
        nonlocal slot, depth, top, below
        while (depth >= 2 or (depth == 1 and has_left)) and hidden(below, top, incoming):
            depth -= 1
            slot -= 1
            top = below
            below = rows[slot - 2] if depth >= 2 else left_zero

    for q in range(n):
        filled, value = classes[q], values[q]
        class_words[q >> 5] |= filled << (q & 31)
        if q > 0 and filled != previous:
            pop_hidden((q, q * q))
            depth = 0
            has_left = True
            left_zero = top = (q - 1, (q - 1) ** 2)
        previous = filled
        if value != NONE:
            incoming = (q, value + q * q)
            pop_hidden(incoming)
            assert slot <= q          # the in-place invariant: a slot never passes the read row
            rows[slot] = incoming
            below, top = top, incoming
            depth += 1
            slot += 1
    stored_total = slot

    # ---- phase 2: lockstep sweep; candidates of a run = left zero, stored sites, right zero
    def load(index):
        return rows[index] if index < stored_total else (NONE, NO_HEIGHT)

    cursor = 0
    pending = load(0)
    winner = right_zero = (0, NO_HEIGHT)
    run_end = 0
    previous = 0
    out = [0] * n
    for q in range(n):
        word = class_words[q >> 5]
        filled = (word >> (q & 31)) & 1
        if q == 0 or filled != previous:
            w = q >> 5
            different = ((~word if filled else word) & 0xFFFFFFFF) & ((0xFFFFFFFF << (q & 31)) & 0xFFFFFFFF)
            while different == 0 and w + 1 < num_words:
                w += 1
                different = (~class_words[w] if filled else class_words[w]) & 0xFFFFFFFF
            run_end = min(n, 32 * w + (different & -different).bit_length() - 1) if different else n
            while pending[0] < q:
                cursor += 1
                pending = load(cursor)
            right_zero = (run_end, run_end ** 2) if run_end < n else (0, NO_HEIGHT)
            winner = (q - 1, (q - 1) ** 2) if q > 0 else (0, NO_HEIGHT)
        previous = filled
        while True:
            from_stack = pending[0] < run_end
            candidate = pending if from_stack else right_zero
            if not (candidate[1] - 2 * candidate[0] * q < winner[1] - 2 * winner[0] * q):
                break
            winner = candidate
            if from_stack:
                cursor += 1
                pending = load(cursor)
            else:
                right_zero = (0, NO_HEIGHT)
        out[q] = NONE if winner[1] == NO_HEIGHT else winner[1] - 2 * winner[0] * q + q * q
    return out


FAR_ZERO_SITE = 46341 + 8192
BLOCKED_HEIGHT = 0x5FFFFFFF
ABSENT_POSITION = 0x4000


def lean_line(classes, values):
    """Model of EnvelopeAxisLeanKernel (csrc/edt_envelope_lean.cuh): one pop loop in phase 1 (the
    incoming site at a class change is the boundary's zero site), the per-word run-end table,
    and a phase 2 that walks only stored sites and takes the right zero site as a min()."""
    n = len(classes)
    num_words = (n + 31) // 32
    rows = [None] * n
    class_words = [0] * num_words
    run_end_after = [None] * num_words
    slot = entries = 0
    left_v = -1
    top = below = (0, 0)
    accumulator = 0
    previous = classes[0]

    for q in range(n):
        filled, value = classes[q], values[q]
        change = filled != previous
        previous = filled
        accumulator = (accumulator >> 1) | (filled << 31)
        # pre-filter: keep the site only if it is strictly below the segment between its two
        # neighbour sites (a neighbour of the other class is a zero site); NONE never filters
        before = 0 if change else (values[q - 1] if q > 0 else NONE)
        if q + 1 < n:
            after = 0 if classes[q + 1] != filled else values[q + 1]
        else:
            after = NONE
        on_hull_locally = 2 * value - 2 - before < after
        finite = value != NONE and (on_hull_locally or q == 0)
        own = (q, value + q * q)
        incoming = (q, q * q) if change else own
        if change or finite:
            while entries >= 2 and hidden(below, top, incoming):
                entries -= 1
                slot -= 1
                top = below
                if entries >= 2:
                    if entries == 2 and left_v >= 0:
                        below = (left_v, left_v * left_v)
                    else:
                        below = rows[slot - 2]
        if change:
            for w in range((left_v + 1) >> 5, q >> 5):
                run_end_after[w] = q
            left_v = q - 1
            top = (left_v, left_v * left_v)
            entries = 1
        if finite:
            assert slot <= q
            rows[slot] = own
            slot += 1
            below, top = top, own
            entries += 1
        if (q & 31) == 31:
            class_words[q >> 5] = accumulator
    if n & 31:
        # arithmetic shift: the class of the last row repeats past the end of the line
        sign = accumulator >> 31
        shift = 32 - (n & 31)
        word = accumulator >> shift
        if sign:
            word |= (0xFFFFFFFF << (32 - shift)) & 0xFFFFFFFF
        class_words[num_words - 1] = word
    for w in range((left_v + 1) >> 5, num_words):
        run_end_after[w] = n
    stored_total = slot

    cursor = 0

    def load_pending():
        return rows[cursor] if cursor < stored_total else (ABSENT_POSITION, NO_HEIGHT)

    pending = load_pending()
    winner = (0, NO_HEIGHT)
    candidate_h = BLOCKED_HEIGHT
    right_v = FAR_ZERO_SITE
    run_end = 0
    previous = 0
    out = [0] * n
    for q in range(n):
        w, b = q >> 5, q & 31
        word = class_words[w]
        filled = (word >> b) & 1
        if filled != previous or q == 0:
            different = ((~word if filled else word) & 0xFFFFFFFF) >> b
            if different:
                run_end = min(n, q + (different & -different).bit_length() - 1)
            else:
                run_end = run_end_after[w]
            while pending[0] < q:
                cursor += 1
                pending = load_pending()
            winner = (q - 1, (q - 1) ** 2) if q > 0 else (0, NO_HEIGHT)
            right_v = run_end if run_end < n else FAR_ZERO_SITE
            candidate_h = pending[1] if pending[0] < run_end else BLOCKED_HEIGHT
            previous = filled
        while candidate_h - 2 * pending[0] * q < winner[1] - 2 * winner[0] * q:
            winner = (pending[0], candidate_h)
            cursor += 1
            pending = load_pending()
            candidate_h = pending[1] if pending[0] < run_end else BLOCKED_HEIGHT
        squared = NONE if winner[1] == NO_HEIGHT else winner[1] - 2 * winner[0] * q + q * q
        out[q] = min(squared, (right_v - q) ** 2, NONE)
    return out


def brute_line(classes, values):
    n = len(classes)
    out = []
    for q in range(n):
        best = NONE
        # the run containing q
        a = q
        while a > 0 and classes[a - 1] == classes[q]:
            a -= 1
        b = q
        while b < n - 1 and classes[b + 1] == classes[q]:
            b += 1
        for v in range(a, b + 1):
            if values[v] != NONE:
                best = min(best, (q - v) ** 2 + values[v])
        if a > 0:
            best = min(best, (q - a + 1) ** 2)
        if b < n - 1:
            best = min(best, (b + 1 - q) ** 2)
        out.append(best)
    return out


def random_line(rng, n):
    flip = rng.choice([0.02, 0.1, 0.5])
    classes, c = [], rng.randint(0, 1)
    for _ in range(n):
        if rng.random() < flip:
            c ^= 1
        classes.append(c)
    scale = rng.choice([3, 30, 400, 5000, 200000])
    none_rate = rng.choice([0.0, 0.2, 0.9])
    values = [NONE if rng.random() < none_rate else rng.randint(1, scale) for _ in range(n)]
    return classes, values


def test_kernel_line_model_matches_brute_force():
    rng = random.Random(1)
    for _ in range(6000):
        n = rng.choice([1, 2, 3, 4, 5, 7, 31, 32, 33, 36, 40, 64, 65, 97])
        classes, values = random_line(rng, n)
        want = brute_line(classes, values)
        assert kernel_line(classes, values) == want, (classes, values)
        assert lean_line(classes, values) == want, (classes, values)


def test_smooth_distance_like_lines():
    # values shaped like real partial distances (smooth bowls), the deep-stack case
    rng = random.Random(2)
    for _ in range(300):
        n = rng.choice([64, 100, 130])
        centre, offset = rng.uniform(0, n), rng.randint(1, 50)
        classes = [1 if abs(i - n / 3) < 4 else 0 for i in range(n)]
        values = [int((i - centre) ** 2) + offset for i in range(n)]
        assert kernel_line(classes, values) == brute_line(classes, values)
        assert lean_line(classes, values) == brute_line(classes, values)



# ------------------------------------------------------------------------------------------------
# Window kernel (csrc/edt_envelope_window.cuh): per voxel a class-agnostic min over the rows
# within R of it plus the nearest opposite-class row from a class-bit window; a result below
# (R + 1)^2 is exact by construction (every row outside the window costs at least (R + 1)^2),
# anything else takes the extended search. Mirrors the kernel's chunking and bit arithmetic.
# ------------------------------------------------------------------------------------------------
def clz32(x):
    return 32 - x.bit_length()


def brev32(x):
    return int(format(x & 0xFFFFFFFF, "032b")[::-1], 2)


SATURATED = 0x7800


DEEPEST_CAP = 168


def window_line(classes, values, radius, step_budget=None, segment_rows=None, deepest=None):
    """Returns (out, steps), or (None, steps) when the line is given up: the joint search ran
    past its budget or its depth cap, or a row that needs it has a saturated minimum.
    segment_rows (a multiple of the radius): the line is cut into segments computed
    independently, as the kernel's grid.y does; a segment that does not start the line reads its
    neighbour's rows.

    Per chunk of R rows, as the kernel: phase A = window minimum + nearest opposite-class row
    inside the window for every row (16-bit results); phase B, only when some row is not
    certified = the rows of the register window beyond each row's own window (class-agnostic,
    plus the nearest opposite-class row of the neighbouring chunks), then rows loaded from memory
    that serve ALL rows of the chunk at once, four per side and round; phase C = emit."""
    n = len(classes)
    chunk = radius
    assert radius % 2 == 0
    far = (radius + 1) ** 2
    steps = 0
    if deepest is None:
        deepest = DEEPEST_CAP

    def u16(x):
        assert 0 <= x < 65536, x
        return x

    out = [0] * n
    mask_r = (1 << radius) - 1
    garbage = random.Random(n)
    if segment_rows is None:
        segment_rows = ((n + chunk - 1) // chunk) * chunk
    assert segment_rows % chunk == 0
    for first_row in range(0, n, segment_rows):
        end_row = min(first_row + segment_rows, n)

        def load_rows(start):
            # rows start .. start + chunk - 1 (load_chunk with an arbitrary first row)
            vals, bits = [], 0
            for i in range(chunk):
                r = start + i
                rc = min(max(r, 0), n - 1)
                vals.append(min(values[rc], SATURATED) if 0 <= r < n else SATURATED)
                bits = ((bits << 1) | classes[rc]) & 0xFFFFFFFF
            return vals, bits

        a, ca = load_rows(first_row - chunk)
        b, cb = load_rows(first_row)
        nx, cn = load_rows(first_row + chunk)
        for base in range(first_row, end_row, chunk):
            composite = (ca << (2 * chunk)) | (cb << chunk) | cn
            rows = a + b + nx  # chunk-relative row r at index r + chunk
            # (bits above 3 R: leftovers of older chunks in the kernel's class word - anything)
            composite |= garbage.getrandbits(8) << (3 * chunk)
            shift = max(3 * chunk - 32, 0)
            boundaries = composite ^ (composite >> 1)
            boundaries_low = boundaries & 0xFFFFFFFF
            boundaries_reversed = brev32((boundaries >> shift) & 0xFFFFFFFF)
            # ---- phase A
            best = [0] * chunk
            for j in range(chunk):
                q = base + j
                if q >= n:
                    break
                acc = 0xFFFF
                for g in range(j >> 1, (j >> 1) + radius + 1):
                    for index in (2 * g, 2 * g + 1):
                        acc = min(acc, u16(rows[index] + (index - chunk - j) ** 2))
                query_class = (composite >> (2 * chunk - 1 - j)) & 1
                assert query_class == classes[q]
                # boundary bits (bit t: the rows at bits t and t + 1 differ); the nearest
                # opposite-class row after / before the row at bit p from one shift and one
                # count of leading zeros per side; min with R = none inside the window
                p = 2 * chunk - 1 - j
                after = clz32((boundaries_low << (32 - p)) & 0xFFFFFFFF)
                before = clz32((boundaries_reversed << (p - shift)) & 0xFFFFFFFF)
                index = min(after, before, radius)
                e_squared = (index + 1) ** 2 if index < radius else 0xFFFF
                best[j] = min(acc, e_squared)
            # ---- phase B
            if max(best) >= far:
                chunk_class = (cb >> (chunk - 1)) & 1
                uniform = cb == (mask_r if chunk_class else 0)
                flip = mask_r if chunk_class else 0
                # nearest opposite-class row of the previous / next chunk (lanes whose chunk is
                # of one class only; 200 = none)
                d_prev = (ca ^ flip) & mask_r
                d_next = (cn ^ flip) & mask_r
                c_below = 200
                c_above = 200
                if uniform and d_prev:
                    c_below = 1 + ((d_prev & -d_prev).bit_length() - 1)
                if uniform and d_next:
                    c_above = 1 + (chunk - 1 - (d_next.bit_length() - 1))
                for j in range(chunk):
                    best[j] = min(best[j], u16((c_below + j) ** 2))
                    best[j] = min(best[j], u16((c_above + chunk - 1 - j) ** 2))
                # class-agnostic candidates: the rows of the register window a row's own window
                # did not reach (pairs of output rows: the low row may see a row twice)
                for k in range(chunk - 1):          # row base - R + k
                    for pair in range(chunk // 2):
                        if 2 * pair + 1 > k:
                            for j in (2 * pair, 2 * pair + 1):
                                best[j] = min(best[j], u16(a[k] + (j + chunk - k) ** 2))
                for m in range(1, chunk):           # row base + R + m
                    for pair in range(chunk // 2):
                        if 2 * pair < m:
                            for j in (2 * pair, 2 * pair + 1):
                                best[j] = min(best[j], u16(nx[m] + (chunk + m - j) ** 2))
                t = 0
                while True:
                    worst = max(best)
                    below_open = base - chunk - 1 - t >= 0
                    above_open = base + 2 * chunk + t <= n - 1
                    if not ((chunk + 1 + t) ** 2 < worst and (below_open or above_open)):
                        break
                    if (step_budget is not None and steps > step_budget) \
                            or chunk + 1 + t + 3 > deepest:
                        return None, steps
                    steps += 4
                    for u in range(4):
                        for row, distance in ((max(base - chunk - 1 - t - u, 0),
                                               lambda j: j + chunk + 1 + t + u),
                                              (min(base + 2 * chunk + t + u, n - 1),
                                               lambda j: 2 * chunk + t + u - j)):
                            height = min(values[row], SATURATED) \
                                if classes[row] == chunk_class else 0
                            for j in range(chunk):
                                best[j] = min(best[j], u16(height + distance(j) ** 2))
                    t += 4
                if max(best) >= SATURATED:
                    return None, steps
            # ---- phase C
            for j in range(chunk):
                if base + j < n:
                    out[base + j] = min(best[j], NONE)
            a, ca = b, cb
            b, cb = nx, cn
            nx, cn = load_rows(base + 2 * chunk)
    return out, steps


def test_window_line_model_matches_brute_force():
    rng = random.Random(3)
    for _ in range(4000):
        n = rng.choice([1, 2, 3, 4, 5, 7, 12, 13, 24, 25, 31, 32, 33, 36, 40, 64, 65, 97])
        classes, values = random_line(rng, n)
        want = brute_line(classes, values)
        for radius in (4, 8, 12, 14):
            got, _ = window_line(classes, values, radius)
            # (None: a saturated window, the line goes to the stack kernel)
            assert got is None or got == want, (radius, classes, values)
            if max(values) < SATURATED:
                assert got == want
            # the same line cut into segments of one and of three chunks
            for segment_chunks in (1, 3):
                cut, _ = window_line(classes, values, radius, segment_rows=segment_chunks * radius)
                assert cut is None or cut == want, (radius, segment_chunks, classes, values)
                assert (cut is None) == (got is None)


def test_window_line_smooth_and_budget():
    rng = random.Random(4)
    solved = 0
    for _ in range(200):
        n = rng.choice([64, 100, 130])
        centre, offset = rng.uniform(0, n), rng.randint(1, 50)
        classes = [1 if abs(i - n / 3) < 4 else 0 for i in range(n)]
        values = [int((i - centre) ** 2) + offset for i in range(n)]
        want = brute_line(classes, values)
        got, steps = window_line(classes, values, 8, deepest=DEEPEST_CAP)
        # (None: some row needs rows further away than the depth cap of the 16-bit search)
        assert got is None or got == want
        solved += got is not None
        # a budget below the steps taken reports the line for the stack kernel instead
        if got is not None and steps > 0:
            assert window_line(classes, values, 8, step_budget=steps - 5,
                               deepest=DEEPEST_CAP)[0] is None
    assert solved > 100
    # one class, no values at all: the window is saturated, the line is given up
    assert window_line([0] * 50, [NONE] * 50, 8)[0] is None


def test_window_line_deep_pockets_stay_inside_sixteen_bits():
    # long lines with pockets up to and beyond the depth cap: the model asserts that every sum
    # of the joint search fits 16 bits; lines within the cap come out exact, deeper ones give up
    rng = random.Random(9)
    solved = gave_up = 0
    for _ in range(60):
        n = rng.choice([300, 420, 520])
        wall = rng.randint(0, 40)
        depth = rng.choice([90, 140, 165, 200, 260])
        # one class; small values near the two ends of a pocket of the given depth, large inside
        values = []
        for i in range(n):
            edge = min(abs(i - wall), abs(i - (wall + 2 * depth)))
            values.append(1 + rng.randint(0, 3) if edge < 2 else 100000 + rng.randint(0, 50))
        classes = [0] * n
        want = brute_line(classes, values)
        for radius in (8, 12):
            got, _ = window_line(classes, values, radius)
            assert got is None or got == want
            solved += got is not None
            gave_up += got is None
            if max(want) < 150 * 150:
                assert got == want, (n, wall, depth, radius)
    assert solved > 20 and gave_up > 5
