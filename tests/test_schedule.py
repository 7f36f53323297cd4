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
