"""-m gpu: layout adapters -- every source/target pair against numpy, and
a packed / scattered batch run through the adapter in front of the Diagonalize path."""
import numpy as np
import pytest

from oracle import binding as ob
import parity

pytestmark = pytest.mark.gpu


def _as_layout(A, layout):
    B, n, _ = A.shape
    iu = np.triu_indices(n)
    if layout == 0:
        return np.ascontiguousarray(A)
    if layout == 1:
        return np.ascontiguousarray(A.transpose(1, 2, 0))
    P = A[:, iu[0], iu[1]]
    return np.ascontiguousarray(P if layout == 2 else P.T)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,B", [(3, 1000), (4, 777), (7, 65), (32, 33), (50, 5)])
def test_convert_all_pairs(n, B, dtype):
    import torch
    import jacobi_pd_b200 as jpd
    A = ob.fill_synthetic(n, B, seed=3 + n, dtype=dtype)          # symmetric
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    for src_layout in (0, 1, 2, 3):
        d = torch.from_numpy(_as_layout(A, src_layout)).cuda()
        for dst_layout in (0, 1, 2, 3):
            out = jpd.convert_layout(d, src_layout, dst_layout, n, B)
            assert np.array_equal(out.cpu().numpy(), _as_layout(A, dst_layout)), (src_layout, dst_layout)
    # scattered matrices through a pointer array (in reverse order, with gaps)
    pool = torch.zeros((2 * B, n, n), dtype=tdt, device="cuda")
    idx = torch.arange(B - 1, -1, -1, device="cuda") * 2
    pool[idx] = torch.from_numpy(A).cuda()
    ptrs = (pool.data_ptr() + idx * (n * n * pool.element_size())).to(torch.int64).contiguous()
    for dst_layout in (0, 1, 2, 3):
        out = jpd.convert_layout(ptrs, jpd.LAYOUT_POINTERS, dst_layout, n, B, dtype=tdt)
        assert np.array_equal(out.cpu().numpy(), _as_layout(A, dst_layout))


def test_packed_input_through_adapter_matches_direct_call():
    import torch
    import jacobi_pd_b200 as jpd
    n, B = 4, 4096
    A = ob.fill_synthetic(n, B, seed=99)
    A_lower_garbage = np.triu(A) + np.tril(np.full_like(A, 99.0), -1)   # the lower triangle is never read
    ref = ob.oracle_diagonalize(A, 1)
    packed = torch.from_numpy(_as_layout(A_lower_garbage, 2)).cuda()
    soa = jpd.convert_layout(packed, jpd.LAYOUT_PACKED_AOS, jpd.LAYOUT_SOA, n, B)
    ev, vec, it = jpd.diagonalize_batched(soa, layout=jpd.LAYOUT_SOA)
    ev0, vec0, it0 = jpd.diagonalize_batched(torch.from_numpy(np.ascontiguousarray(A.transpose(1, 2, 0))).cuda(),
                                             layout=jpd.LAYOUT_SOA)
    assert torch.equal(ev, ev0) and torch.equal(vec, vec0) and torch.equal(it, it0)
    vec_aos = jpd.convert_layout(vec, jpd.LAYOUT_SOA, jpd.LAYOUT_AOS, n, B)   # eigenvector rows back to AoS
    parity.full_check(A, (ev.cpu().numpy().T.copy(), vec_aos.cpu().numpy(), it.cpu().numpy()), ref, 1)
