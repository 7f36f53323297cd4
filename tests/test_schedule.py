"""CPU checks of the static schedules of tiers 2/3 (host restatement of the index logic of
jacobi_pd_b200/csrc/jpd_warp.cuh: odd-even ordering, build_schedule; jpd_cta.cuh:
cta_build_schedule).  No GPU, no oracle: pure combinatorics the kernels rely on."""
import itertools

import pytest


def odd_even_sweep(m):
    """Pairs met in one sweep of m steps (patterns A, B alternating, positions exchanged)."""
    pos = list(range(m))
    met = []
    for step in range(m):
        for p in range(step % 2, m - 1, 2):
            met.append((min(pos[p], pos[p + 1]), max(pos[p], pos[p + 1])))
            pos[p], pos[p + 1] = pos[p + 1], pos[p]
    return met, pos


@pytest.mark.parametrize("m", list(range(2, 130, 2)))
def test_odd_even_ordering_meets_every_pair_once_and_reverses(m):
    met, pos = odd_even_sweep(m)
    assert sorted(met) == sorted(itertools.combinations(range(m), 2))   # each pair exactly once
    assert pos == list(range(m - 1, -1, -1))                             # order reversed (odd-n padding rule)


def shape(NP):            # WarpShape<NP> / CtaShape<NPB>
    if NP <= 32:
        kld = 9 if NP == 8 else NP + 3
        classes = min(NP, 16)
        return dict(kld=kld, kh=NP // 2, mul=(2 * kld) & 15, classes=classes, halves=NP // classes,
                    kit={32: 4, 16: 2, 8: 1}[NP], lanes=NP)
    threads = NP * NP // 32
    return dict(kld=NP + 3, kh=NP // 2, mul=(2 * (NP + 3)) & 15, classes=16, halves=threads // 16, kit=4, lanes=threads)


def build_schedule(NP, npr, off):
    sh = shape(NP)
    tbl = {}
    for cl in range(sh["classes"]):
        slot = 0
        for K in range(npr):
            L = (cl - sh["mul"] * K) % sh["classes"]
            while L < npr:
                if L > K:
                    idx = (slot // sh["halves"]) * sh["lanes"] + (slot % sh["halves"]) * sh["classes"] + cl
                    assert idx < sh["kit"] * sh["lanes"], "schedule table overflow: a block would be dropped"
                    assert idx not in tbl
                    addr = (2 * K + off) * sh["kld"] + ((2 * L + off) >> 1) + ((2 * L + off) & 1) * sh["kh"]
                    tbl[idx] = (K, L, addr)
                    slot += 1
                L += sh["classes"] if NP > 32 else 10 ** 9
    return tbl, sh


@pytest.mark.parametrize("NP", [8, 16, 32, 64, 96, 128])
def test_schedule_tables_cover_every_block_once_without_bank_conflicts(NP):
    lo = {8: 7, 16: 9, 32: 17, 64: 33, 96: 65, 128: 97}[NP]
    for n in range(lo, NP + 1):
        m = n + (n & 1)
        for off, npr in ((0, m // 2), (1, m // 2 - 1)):
            tbl, sh = build_schedule(NP, npr, off)
            blocks = sorted((K, L) for K, L, _ in tbl.values())
            assert blocks == sorted(itertools.combinations(range(npr), 2))
            assert all(addr < 65536 and K < 256 and L < 256 for K, L, addr in tbl.values())
            if NP < 16:
                continue   # 8-lane groups share a half warp with another matrix: not conflict-free by design
            # every half warp (16 consecutive lanes of one iteration) hits 16 different 8-byte banks
            o01 = sh["kh"] if off == 0 else 1 - sh["kh"]
            for it in range(sh["kit"]):
                for hw in range(sh["lanes"] // 16):
                    base = it * sh["lanes"] + hw * 16
                    ads = [tbl[base + l][2] for l in range(16) if base + l in tbl]
                    for delta in (0, o01, sh["kld"], sh["kld"] + o01):
                        banks = [(a + delta) % 16 for a in ads]
                        assert len(set(banks)) == len(banks)


# ---- tier 4 (jpd_cluster.cuh): snake ownership of 16-row groups, per-CTA schedule tables
def cl_owner(r):
    g = (r >> 4) & 3
    return 3 - g if (r >> 6) & 1 else g


def cl_lrow(r):
    return ((r >> 6) << 4) | (r & 15)


def cl_grow(lr, rank):
    sg = lr >> 4
    return 64 * sg + 16 * ((3 - rank) if sg & 1 else rank) + (lr & 15)


def cl_pair_of(t, rank):
    sg = t >> 3
    return 8 * (4 * sg + ((3 - rank) if sg & 1 else rank)) + (t & 7)


def test_cluster_ownership_maps_are_consistent():
    for rank in range(4):
        rows = [cl_grow(lr, rank) for lr in range(64)]
        assert all(cl_owner(r) == rank and cl_lrow(r) == lr for lr, r in enumerate(rows))
    assert sorted(cl_grow(lr, rank) for rank in range(4) for lr in range(64)) == list(range(256))
    for off in (0, 1):      # every pair is handled by exactly one pivot quad, in the CTA owning its first row
        pairs = sorted((cl_pair_of(t, rank), rank) for rank in range(4) for t in range(32))
        assert [p for p, _ in pairs] == list(range(128))
        assert all(cl_owner(2 * j + off) == rank for j, rank in pairs if 2 * j + off < 256)


@pytest.mark.parametrize("n", [129, 130, 160, 200, 255, 256])
def test_cluster_schedule_tables(n):
    m = n + (n & 1)
    for off, npr in ((0, m // 2), (1, m // 2 - 1)):
        seen = []
        for rank in range(4):
            used = set()
            for cl in range(16):
                slot = 0
                for K in range(npr):
                    if cl_owner(2 * K + off) != rank:
                        continue
                    L = (cl - 6 * (K & 7)) % 16
                    while L < npr:
                        if L > K:
                            idx = (slot >> 5) * 512 + (slot & 31) * 16 + cl
                            assert idx < 4 * 512, "schedule table overflow"
                            assert idx not in used
                            used.add(idx)
                            seen.append((K, L))
                            slot += 1
                        L += 16
        assert sorted(seen) == sorted(itertools.combinations(range(npr), 2))
    if n == 256:   # perfectly balanced at the design size
        for rank in range(4):
            cnt = sum(1 for K in range(128) if cl_owner(2 * K) == rank for L in range(K + 1, 128))
            assert cnt == 2032


# ---- tier 3, block-ordering kernel (jpd_cta4.cuh): block odd-even sweep, tile storage, tile tables
def block_sweep(m):
    """Pairs met in one sweep of the block ordering (m a multiple of 4): the step inside the blocks,
    then m/2 block steps A B A B ...; a block step on positions q0..q3 rotates (q0,q2)(q1,q3) with
    exchange and (q0,q3)(q1,q2) without."""
    pos = list(range(m))
    met = []

    def rot(p, q, exchange):
        met.append((min(pos[p], pos[q]), max(pos[p], pos[q])))
        if exchange:
            pos[p], pos[q] = pos[q], pos[p]

    for p in range(0, m, 2):
        rot(p, p + 1, False)
    nb = m // 2
    for bs in range(nb):
        for J in range(bs % 2, nb - 1, 2):
            q0 = 2 * J
            rot(q0, q0 + 2, True); rot(q0 + 1, q0 + 3, True)
            rot(q0, q0 + 3, False); rot(q0 + 1, q0 + 2, False)
    return met, pos


@pytest.mark.parametrize("m", list(range(4, 132, 4)))
def test_block_ordering_meets_every_pair_once_and_reverses_the_blocks(m):
    met, pos = block_sweep(m)
    assert sorted(met) == sorted(itertools.combinations(range(m), 2))
    nb = m // 2
    # the kernel's orig_index(): position p holds index 2 (nb-1 - p/2) + (p&1) after an odd number of sweeps
    assert pos == [2 * (nb - 1 - (p >> 1)) + (p & 1) for p in range(m)]


def cta4_shape(NPB):
    kbp = NPB // 4
    return dict(kbp=kbp, kw=((kbp + 15) // 16) * 16 + 1, threads=NPB * NPB // 32, kit=2)


def tile4_off(off, w, r, c):
    if off == 0:
        return 4 * r + c
    return (w * 17 if r >= 2 else 0) + (17 if c >= 2 else 0) + 4 * ((r + 2) & 3) + ((c + 2) & 3)


@pytest.mark.parametrize("NPB", [64, 96, 128])
def test_cta4_tile_storage_and_tables(NPB):
    sh = cta4_shape(NPB)
    w = sh["kw"]
    addr = lambda i, j: ((i >> 2) * w + (j >> 2)) * 17 + 4 * (i & 3) + (j & 3)
    for n in range(NPB - 31, NPB + 1):
        m = (n + 3) & ~3
        nbp = m // 4
        # storage: distinct offsets for the upper triangle, inside the reserved area
        offs = [addr(i, j) for i in range(m) for j in range(i, m)]
        assert len(set(offs)) == len(offs) and max(offs) < sh["kbp"] * w * 17
        for pat, np_ in ((0, nbp), (1, nbp - 1)):
            off = 2 * pat
            tbl = {}
            for cl in range(16):
                slot = 0
                for P in range(np_):
                    Q = P + 1 + ((cl - 2 * P - 1) & 15)
                    while Q < np_:
                        idx = (slot // (sh["threads"] // 16)) * sh["threads"] + (slot % (sh["threads"] // 16)) * 16 + cl
                        assert idx < sh["kit"] * sh["threads"], "tile table overflow: a tile would be dropped"
                        assert idx not in tbl
                        tbl[idx] = (P, Q)
                        slot += 1
                        Q += 16
            assert sorted(tbl.values()) == sorted(itertools.combinations(range(np_), 2))
            # the 16 elements of a pattern-OFF tile are the elements the ordering says, in the stored triangle
            for P, Q in tbl.values():
                for r in range(4):
                    for c in range(4):
                        assert (P * w + Q) * 17 + tile4_off(off, w, r, c) == addr(4 * P + off + r, 4 * Q + off + c)
            # every half warp of an iteration touches 16 different 8-byte banks, whatever the element
            for it in range(sh["kit"]):
                for hw in range(sh["threads"] // 16):
                    base = it * sh["threads"] + hw * 16
                    ids = [tbl[base + l][0] * w + tbl[base + l][1] for l in range(16) if base + l in tbl]
                    banks = [(i * 17) % 16 for i in ids]
                    assert len(set(banks)) == len(banks)


# ---- tier 4, block-ordering kernel (jpd_cluster4.cuh): tile rows dealt in groups of eight, tile tables with the
# straddling rows first, the two record ranges a CTA sends, the strips of pattern B
def _cl4_owner_row(P):
    g = P >> 3
    return g if g < 4 else 7 - g


def _cl4_local_row(P):
    return (0 if (P >> 3) < 4 else 8) + (P & 7)


def _cl4_global_row(l, rank):
    return 8 * rank + l if l < 8 else 8 * (7 - rank) + (l - 8)


def test_cluster4_ownership_and_record_ranges():
    for P in range(64):
        r, l = _cl4_owner_row(P), _cl4_local_row(P)
        assert 0 <= r < 4 and 0 <= l < 16 and _cl4_global_row(l, r) == P
    # the two contiguous ranges of pivot records a CTA sends (cl4_front): every block pair exactly once over the CTAs,
    # each range = the pairs of pivot threads 0..7 / 8..15 of that CTA
    seen = []
    for rank in range(4):
        for J0 in (8 * rank, 8 * (7 - rank)):
            rng = list(range(J0, J0 + 8))
            assert all(_cl4_owner_row(J) == rank for J in rng)
            seen += rng
        assert sorted(_cl4_global_row(l, rank) for l in range(16)) == sorted(seen[-16:])
    assert sorted(seen) == list(range(64))
    # the scales of position p live with the owner of tile row p / 4: a pattern-B pair J reads positions 4J+2 .. 4J+5,
    # the last two of which are remote exactly when tile row J + 1 belongs to another CTA
    for J in range(63):
        remote = _cl4_owner_row(J + 1) != _cl4_owner_row(J)
        owners = [_cl4_owner_row(p >> 2) for p in range(4 * J + 2, 4 * J + 6)]
        assert owners[:2] == [_cl4_owner_row(J)] * 2 and (owners[2] != owners[0]) == remote and owners[2] == owners[3]


@pytest.mark.parametrize("n", [129, 130, 176, 200, 253, 256])
def test_cluster4_tile_tables_and_strips(n):
    m = (n + 3) & ~3
    nbp = m // 4
    W, threads = 65, 512
    for pat, np_ in ((0, nbp), (1, nbp - 1)):
        all_tiles = []
        for rank in range(4):
            tbl, overflow = {}, []
            for cl in range(16):
                slot = 0
                for ll in range(32):                      # straddling tile rows first (jpd_cluster4.cuh, table builder)
                    l = ll % 16
                    P = _cl4_global_row(l, rank)
                    if P >= np_:
                        continue
                    remote = pat == 1 and _cl4_owner_row(P + 1) != rank
                    if (ll < 16) != remote:
                        continue
                    Q = P + 1 + ((cl - l - P - 1) & 15)
                    while Q < np_:
                        assert (l + Q) % 16 == cl
                        if slot < threads // 16:
                            tbl[slot * 16 + cl] = (P, Q, remote)
                            slot += 1
                        else:
                            overflow.append((P, Q, remote))
                        Q += 16
            free = [t for t in range(threads) if t not in tbl]
            assert len(overflow) <= len(free), "a tile would be dropped"
            for t, ent in zip(free, overflow):
                tbl[t] = ent
            all_tiles += [(P, Q) for P, Q, _ in tbl.values()]
            # straddling tiles sit in the first warps: with at most four per bank class they stay below thread 64 + 16
            strad = [t for t, (_, _, rem) in tbl.items() if rem]
            if strad:
                assert max(strad) < 16 * (max(sum(1 for t in strad if t % 16 == cl) for cl in range(16)))
            # a half warp touches 16 different 8-byte banks (class-assigned entries)
            for hw in range(threads // 16):
                ids = [_cl4_local_row(tbl[hw * 16 + k][0]) * W + tbl[hw * 16 + k][1] for k in range(16)
                       if hw * 16 + k in tbl and (_cl4_local_row(tbl[hw * 16 + k][0]) + tbl[hw * 16 + k][1]) % 16 == k]
                assert len({(i * 17) % 16 for i in ids}) == len(ids)
        assert sorted(all_tiles) == sorted(itertools.combinations(range(np_), 2))
    # pattern-B strips: 64 top strips on lanes 0..3 of the 16 warps, 16 right strips on the last half warp
    strips = {}
    for tid in range(threads):
        lane, warp = tid & 31, tid >> 5
        r = warp * 4 + lane if lane < 4 else (64 + tid - (threads - 16) if tid >= threads - 16 else None)
        if r is not None:
            assert r not in strips
            strips[r] = tid
    assert sorted(strips) == list(range(80))
