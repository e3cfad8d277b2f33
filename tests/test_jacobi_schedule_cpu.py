"""CPU: the pair ordering of the cluster-resident LRKD eigensolver (deltakd_b200/csrc/lrkd.cu, section 3b), restated in
Python from the kernel's index arithmetic: every column pair must be rotated exactly once per sweep — across the 16 CTAs
(recursive halving over sub-rings), inside a CTA (three slot-0 columns per warp against walking slot-1 triples) and inside a
group (tournament over the 4 column triples) — and sweeps must chain without returning the columns home."""
import itertools

CS, G, C = 16, 12, 3          # CTAs per cluster, columns per group, columns of each slot per warp
W = G // C                    # warps per CTA
T = 2 * CS - 1                # group rounds per sweep


def destinations(t, rank):
    """(cta, slot) that receive the content of slot 0 and slot 1 of CTA `rank` after group round t (kernel: dcta0 / dslot0 / dcta1 / dslot1)."""
    R, r = CS, t
    while r >= R:
        r -= R
        R >>= 1
    i = rank & (R - 1)
    base = rank - i
    d0, d1 = (rank, 0), (rank, 1)
    if R > 1:
        if r < R - 1:
            d1 = (base + ((i + 1) & (R - 1)), 1)
        elif i < R // 2:
            d1 = (rank + R // 2, 0)
        else:
            d0 = (rank - R // 2, 1)
    return d0, d1


def cross_steps():
    """Pairs (slot-0 column, slot-1 column) of one group round, step by step, for all warps of a CTA."""
    steps = []
    for m in range(W):                    # macro rounds
        for s in range(C):                # steps of a macro round: pair i = (x_i, y_(i+s) mod C)
            pairs = []
            for w in range(W):
                j = (w + m) & (W - 1)
                pairs += [(C * w + i, C * j + (i + s) % C) for i in range(C)]
            steps.append(pairs)
    return steps


def inside_steps():
    """Pairs of columns of ONE group rotated in the last group round of a sweep (two warps per group)."""
    steps = []
    for rr in range(3):
        per_warp = []
        for v in range(2):
            vj = 3 - v
            sa = 0 if v == 0 else 1 + (v - 1 + rr) % 3
            sb = 1 + (vj - 1 + rr) % 3
            cols = [C * sa + c for c in range(C)] + [C * sb + c for c in range(C)]
            st = []
            if rr == 0:                   # pairs inside each of the two triples: (0,1) (0,2) (1,2), one pair per triple and step
                for a, b in ((0, 1), (0, 2), (1, 2)):
                    st.append([(cols[a], cols[b]), (cols[C + a], cols[C + b])])
            for s in range(C):
                st.append([(cols[i], cols[C + (i + s) % C]) for i in range(C)])
            per_warp.append(st)
        for k in range(len(per_warp[0])):
            steps.append(per_warp[0][k] + per_warp[1][k])
    return steps


def test_cross_pairs_of_a_group_round_cover_the_12_x_12_block_once():
    steps = cross_steps()
    assert len(steps) == G
    seen = set()
    for pairs in steps:
        xs, ys = [p[0] for p in pairs], [p[1] for p in pairs]
        assert len(set(xs)) == G and len(set(ys)) == G      # a step touches every column once: its pairs are independent
        seen |= set(pairs)
    assert seen == set(itertools.product(range(G), range(G)))


def test_inside_pairs_cover_a_group_once():
    steps = inside_steps()
    assert len(steps) == 12
    seen = []
    for pairs in steps:
        cols = [c for p in pairs for c in p]
        assert len(cols) == len(set(cols))                   # independent pairs
        seen += [tuple(sorted(p)) for p in pairs]
    assert sorted(seen) == sorted(itertools.combinations(range(G), 2))


def test_group_schedule_meets_every_pair_of_groups_once_per_sweep_and_chains():
    slots = [[2 * c, 2 * c + 1] for c in range(CS)]          # group ids
    for sweep in range(3):
        met = set()
        for t in range(T):
            for c in range(CS):
                key = tuple(sorted(slots[c]))
                assert key not in met
                met.add(key)
            new = [[None, None] for _ in range(CS)]
            for c in range(CS):
                d0, d1 = destinations(t, c)
                assert new[d0[0]][d0[1]] is None and new[d1[0]][d1[1]] is None
                new[d0[0]][d0[1]] = slots[c][0]
                new[d1[0]][d1[1]] = slots[c][1]
            # only ONE group per CTA crosses the cluster network, except at the four phase changes
            moved = sum(1 for c in range(CS) for d in destinations(t, c) if d[0] != c)
            assert moved <= CS
            slots = new
        assert len(met) == 2 * CS * (2 * CS - 1) // 2


def test_sequential_depth_of_a_sweep():
    # 31 group rounds x 12 steps + 12 steps inside the groups: 384 dependent pair steps per sweep for 384 columns
    assert T * len(cross_steps()) + len(inside_steps()) == 384
