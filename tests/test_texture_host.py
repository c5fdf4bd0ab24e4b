"""Host side of the Leung-Malik tensor-core path (no GPU): the operand layout `texture.bank_operand_layout` hands to `isb_lm_texture`
must be what the kernel's descriptors address -- K-major core matrices of 8 filters x 4 taps, value and remainder rounded to tf32, the
column (filter) order the epilogue's battery table assumes (csrc/lm_texture.cu `lm_batt_of_col`)."""
import numpy as np
import pytest


def _batt_of_col(col, gs, ng, ns):
    """restatement of lm_batt_of_col: oriented groups first (edge s0 | bar s0 | edge s1 | ...), then Gauss / LoG / LoG2 per sigma"""
    if col < gs * ng:
        return 5 * ((col // gs) // 2) + ((col // gs) % 2)
    rest = col - gs * ng
    return 5 * (rest // 3) + 2 + rest % 3 if rest < ns else -1


@pytest.mark.parametrize('bank', ['normal', 'short'])
def test_operand_layout_of_the_filter_bank(bank):
    from pyimsegm_b200 import texture
    from pyimsegm_b200.descriptors import SHORT_FILTERS_SIGMAS
    names, w_tc, NP, orient, n_batt = texture.bank_operand_layout(bank)
    filters, ref_names = (texture.create_filter_bank_lm_2d(sigmas=SHORT_FILTERS_SIGMAS, nb_orient=4) if bank == 'short'
                          else texture.create_filter_bank_lm_2d())
    assert names == ref_names and n_batt == len(filters) == (15 if bank == 'short' else 20)
    assert w_tc.dtype == np.float32 and w_tc.shape == (33, 2, 10, NP // 8, 8, 4) and w_tc.flags.c_contiguous
    # both halves are representable in tf32 (10 explicit mantissa bits)
    assert not np.any(w_tc.view(np.uint32) & np.uint32(0x1FFF))
    # back to [kernel row][tap][filter]
    dense = w_tc.transpose(0, 1, 2, 5, 3, 4).reshape(33, 2, 40, NP)
    assert not dense[:, :, 33:, :].any()                                  # taps 33..39 are padding
    total = dense[:, 0].astype(np.float64) + dense[:, 1].astype(np.float64)
    n_sig = n_batt // 5
    gs, ng, ns = orient, 2 * n_sig, 3 * n_sig
    seen = {}
    for col in range(NP):
        b = _batt_of_col(col, gs, ng, ns)
        if b < 0:
            assert not dense[:, :, :, col].any()                          # padding filters
            continue
        k = seen.get(b, 0)
        seen[b] = k + 1
        kern = np.asarray(filters[b][k], dtype=np.float64)[::-1, ::-1]    # correlation with the flipped kernel == ndimage.convolve
        # value + remainder reproduce the float32 kernel to two tf32 roundings (relative 2^-22 of the entry)
        k32 = kern.astype(np.float32).astype(np.float64)
        np.testing.assert_allclose(total[:, :33, col], k32, rtol=0, atol=np.abs(k32).max() * 2.0 ** -21)
    assert seen == {b: len(filters[b]) for b in range(n_batt)}            # every kernel of every battery exactly once, in order
