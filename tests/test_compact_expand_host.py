"""tkz_compact_expand is host code (the replay of Encoding.fromTokens + Encoding.pad, src/encoding.zig:246-294, 385-463, over
the compact result of the host-buffer call): exercised here without a GPU on hand-made compact results -- packed and 32-bit
offsets, u16 and u32 ids, padding left / right, and the side list of wide tokens (0xFFFF in the packed array)."""
import ctypes as C

import numpy as np
import pytest

import tokzig_b200 as tz


def make_result(kept_counts, ids, offs, pad=None, ids16=True, packed=True, wide_from=256):
    """a tkz_compact_result over numpy arrays (kept alive by the returned tuple)"""
    r = tz.CompactResult()
    dko = np.zeros(len(kept_counts) + 1, np.uint64)
    np.cumsum(kept_counts, out=dko[1:])
    ids = np.asarray(ids, np.uint32)
    offs = np.asarray(offs, np.uint32).reshape(-1, 2)
    keep = [dko]
    r.n_docs, r.n_kept, r.n_real_tokens = len(kept_counts), len(ids), len(ids)
    r.doc_kept_off = dko.ctypes.data
    if ids16:
        a = ids.astype(np.uint16); keep.append(a); r.ids16 = a.ctypes.data
    else:
        a = ids.copy(); keep.append(a); r.ids = a.ctypes.data
    if packed:
        p16 = np.zeros(len(ids), np.uint16)
        wide = []
        for k, (s, e) in enumerate(offs.tolist()):
            if e >= wide_from:
                p16[k] = 0xFFFF
                wide.append([k & 0xFFFFFFFF, k >> 32, s, e])
            else:
                p16[k] = s | (e << 8)
        w = np.asarray(wide, np.uint32).reshape(-1, 4)
        keep += [p16, w]
        r.offsets_packed = p16.ctypes.data
        r.n_wide = len(w)
        r.wide_tokens = w.ctypes.data if len(w) else None
    else:
        o = offs.copy(); keep.append(o); r.offsets = o.ctypes.data
    if pad is not None:
        r.params.has_padding, r.params.pad_length = 1, pad["length"]
        r.params.pad_id, r.params.pad_type_id, r.params.pad_left = pad["pad_id"], pad["pad_type_id"], 1 if pad["left"] else 0
    return r, keep


def expected(kept_counts, ids, offs, pad, d0, d1):
    ids = np.asarray(ids, np.uint32); offs = np.asarray(offs, np.uint32).reshape(-1, 2)
    dko = np.concatenate([[0], np.cumsum(kept_counts)])
    o_ids, o_off, o_attn, o_typ, o_sp, dto = [], [], [], [], [], [0]
    for d in range(d0, d1):
        k0, k1 = int(dko[d]), int(dko[d + 1])
        n = k1 - k0
        npad = max(0, pad["length"] - n) if pad else 0
        real = (ids[k0:k1].tolist(), offs[k0:k1].tolist(), [1] * n, [0] * n, [0] * n)
        padv = ([pad["pad_id"]] * npad, [[0, 0]] * npad, [0] * npad, [pad["pad_type_id"]] * npad, [1] * npad) if pad else ([], [], [], [], [])
        first, second = (padv, real) if (pad and pad["left"]) else (real, padv)
        for dst, a, b in zip((o_ids, o_off, o_attn, o_typ, o_sp), first, second):
            dst += a; dst += b
        dto.append(len(o_ids))
    return dto, o_ids, o_off, o_attn, o_typ, o_sp


@pytest.mark.parametrize("ids16", [True, False])
@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("pad", [None, {"length": 6, "pad_id": 9, "pad_type_id": 3, "left": False}, {"length": 6, "pad_id": 9, "pad_type_id": 3, "left": True}])
def test_expand_replays_from_tokens_and_pad(ids16, packed, pad):
    rng = np.random.default_rng(5)
    kept = [3, 0, 7, 1, 6, 0, 2]
    n = sum(kept)
    ids = rng.integers(0, 60000, n)
    starts = rng.integers(0, 200, n)
    ends = starts + rng.integers(1, 50, n)
    for k in (2, 3, 9, 18):                                   # tokens of pre-tokens of 256+ bytes
        starts[k] = rng.integers(0, 5000); ends[k] = starts[k] + rng.integers(256, 70000)
    ends[5] = 255; starts[5] = 254                             # the largest pair that still fits the u16
    offs = np.stack([starts, ends], axis=1)
    r, keep = make_result(kept, ids, offs, pad, ids16, packed)
    assert (r.n_wide == 4) == packed
    for d0, d1 in ((0, len(kept)), (2, 3), (3, 7), (4, 5), (1, 2), (6, 7)):
        got = tz.expand_compact(r, d0, d1)
        dto, e_ids, e_off, e_attn, e_typ, e_sp = expected(kept, ids, offs, pad, d0, d1)
        assert got.doc_tok_off.tolist() == dto
        assert got.ids.tolist() == e_ids
        assert got.offsets.tolist() == e_off
        assert got.attention_mask.tolist() == e_attn and got.type_ids.tolist() == e_typ and got.special_tokens_mask.tolist() == e_sp


def test_a_missing_wide_record_is_an_error_not_a_guess():
    kept = [2, 2]
    r, keep = make_result(kept, [1, 2, 3, 4], [[0, 3], [0, 300], [1, 2], [5, 400]])
    assert r.n_wide == 2
    r.n_wide = 1                                                # the record of kept token 3 is gone
    with pytest.raises(tz.TokzigError):
        tz.expand_compact(r, 0, 2)
    assert tz.expand_compact(r, 0, 1).offsets.tolist() == [[0, 3], [0, 300]]
