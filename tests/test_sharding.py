"""Multi-GPU host logic on CPU: byte-balanced document sharding (no collective on the data path) and, with two gloo
ranks, that concatenating the per-rank encodings in rank order reproduces the single-process result.  The per-rank
encode is stood in for by the oracle here (this container has no GPU); the GPU run of the same property is
tests/test_gpu_parity.py::test_batch_split_invariance_and_roundtrip_property."""
import os
import socket
import sys

import numpy as np
import pytest

import tokzig_b200 as tz
from tools import corpus


def test_shard_bounds_properties():
    rng = np.random.default_rng(0)
    for _ in range(50):
        nd = int(rng.integers(0, 200))
        lens = rng.integers(0, 5000, size=nd) * (rng.random(nd) < 0.9)
        off = np.zeros(nd + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        for world in (1, 2, 3, 8):
            b = tz.shard_bounds(off, world)
            assert b[0] == 0 and b[-1] == nd and np.all(np.diff(b) >= 0)
            if nd and off[-1] > 0:
                sizes = off[b[1:]].astype(np.int64) - off[b[:-1]].astype(np.int64)
                assert sizes.sum() == int(off[-1])
                # no shard is further from its target cut than the largest document
                assert np.abs(sizes - int(off[-1]) / world).max() <= lens.max() + 1


def test_skewed_corpus_balance():
    text, off = corpus.generate("c2", 8 << 20, 5)
    b = tz.shard_bounds(off, 8)
    sizes = off[b[1:]].astype(np.int64) - off[b[:-1]].astype(np.int64)
    assert sizes.max() / sizes.mean() < 1.05


def test_document_costs_match_a_brute_force_and_balance_the_skewed_corpus():
    """bytes are not equally expensive on the c5 corpus (unbroken words of KiBs..MiBs go through the block / grid
    kernels at 20-36x the time per byte): the shard cut by modelled cost is never worse than the cut by bytes."""
    import re
    text, off = corpus.generate("c5", 64 << 20, 11)
    cost = tz.document_costs(text, off)
    b = text.tobytes()
    for d in list(range(0, len(off) - 1, 7))[:120]:
        doc = b[int(off[d]):int(off[d + 1])]
        want = float(len(doc))
        for m in re.finditer(rb"[^ \t\n\r]+", doc):
            n = len(m.group(0))
            want += (tz.COST_HUGE_WORD - 1) * n if n > 12288 else ((tz.COST_LONG_WORD - 1) * n if n > 64 else 0.0)
        assert abs(cost[d] - want) < 1e-6 * max(want, 1.0), d
    whole = tz.document_costs(text, off, b"")                      # no pre-tokenizer: the document is the word
    lens = np.diff(off.astype(np.int64))
    assert np.allclose(whole, lens * np.where(lens > 12288, tz.COST_HUGE_WORD, np.where(lens > 64, tz.COST_LONG_WORD, 1.0)))
    acc = np.concatenate([[0.0], np.cumsum(cost)])
    imb = {}
    for world in (2, 4, 8):
        for name, c in (("bytes", None), ("cost", cost)):
            bb = tz.shard_bounds(off, world, c)
            assert bb[0] == 0 and bb[-1] == len(off) - 1 and np.all(np.diff(bb) >= 0)
            per = acc[bb[1:]] - acc[bb[:-1]]
            imb[name, world] = per.max() / per.mean()             # slowest shard over the mean, in modelled time
    # (contiguous cuts at document boundaries: a single 4 MiB word is a floor neither cut can go below)
    assert imb["cost", 4] < 0.9 * imb["bytes", 4] and imb["cost", 8] < 0.8 * imb["bytes", 8], imb
    assert imb["cost", 2] <= imb["bytes", 2] + 0.01, imb
    # an ordinary corpus: cost == bytes up to the rare long word, the cuts barely move
    text2, off2 = corpus.generate("c2", 8 << 20, 5)
    c2 = tz.document_costs(text2, off2)
    assert c2.sum() < 1.05 * int(off2[-1])
    assert np.abs(tz.shard_bounds(off2, 8, c2) - tz.shard_bounds(off2, 8)).max() <= max(8, (len(off2) - 1) // 50)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as orc
    from tools import tokenizers_io
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text, off = corpus.generate("c2", 1 << 20, 42)
    o = orc.OracleTokenizer.from_json(tokenizers_io.tokenizer_json("gpt2_whitespace"))
    t_s, off_s, (lo, hi) = tz.shard(text, off, rank, world)
    enc = o.encode_packed(t_s, off_s)
    # "results are gathered to host only for verification": gather object lists on rank 0
    gathered = [None] * world
    dist.gather_object((lo, hi, enc.ids, enc.offsets, enc.doc_tok_off), gathered if rank == 0 else None, dst=0)
    # timing plumbing used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        full = o.encode_packed(text, off)
        ids = np.concatenate([g[2] for g in gathered])
        offs = np.concatenate([g[3] for g in gathered])
        dto, acc = [np.zeros(1, np.uint64)], 0
        for g in gathered:
            dto.append(g[4][1:] + np.uint64(acc)); acc += int(g[4][-1])
        ok = (np.array_equal(ids, full.ids) and np.array_equal(offs, full.offsets) and np.array_equal(np.concatenate(dto), full.doc_tok_off)
              and gathered[0][0] == 0 and gathered[-1][1] == len(off) - 1 and all(gathered[i][1] == gathered[i + 1][0] for i in range(world - 1)))
        q.put(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_encode_matches_single():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_c_shard_cut_and_cost_model_equal_the_numpy_statement():
    """tkzm_shard_bounds / tkzm_document_costs (the product's multi-GPU cut, C++) against the numpy statements above: random
    documents incl. empty ones, long words across the 64 / 12288-byte thresholds, with and without a pre-tokenizer."""
    import tokzig_b200 as tz
    rng = np.random.default_rng(5)
    for trial in range(6):
        nd = int(rng.integers(1, 4000))
        lens = rng.integers(0, 200, nd).astype(np.uint64)
        lens[rng.random(nd) < 0.05] = 0
        big = rng.random(nd) < 0.01
        lens[big] = rng.integers(1000, 40000, int(big.sum()))
        off = np.zeros(nd + 1, np.uint64); np.cumsum(lens, out=off[1:])
        text = rng.choice(np.frombuffer(b"abcdefghij \n\t\r", np.uint8), int(off[-1]), p=[0.09] * 10 + [0.07, 0.01, 0.01, 0.01]).astype(np.uint8)
        for d in np.flatnonzero(big)[::2]:                      # unbroken runs inside some long documents
            a, b = int(off[d]), int(off[d + 1])
            text[a + 10: b - 10] = ord("x")
        raw_class = np.zeros(256, np.uint8); raw_class[list(b" \t\n\r")] = 1
        for rc_, delim in ((raw_class, b" \t\n\r"), (None, b"")):
            ref = tz.document_costs(text, off, delimiters=delim)
            got = tz.document_costs_c(text, off, rc_, threads=[0, 1, 3][trial % 3])
            assert np.allclose(got, ref), f"trial {trial} delim {delim!r}"
            for n in (1, 2, 3, 8):
                assert np.array_equal(tz.shard_bounds_c(off, n), tz.shard_bounds(off, n))
                assert np.array_equal(tz.shard_bounds_c(off, n, got), tz.shard_bounds(off, n, ref))
