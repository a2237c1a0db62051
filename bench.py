#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric (input GB/s + tokens/s of batch encode) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2b|c2a|c3|c4a|c4b|c5a|c5b] [--size-mib M]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      the reference's algorithm (C restatement, oracle/) on the host cores

A step = one pass of the encode hot path over one synthetic batch (default: configs[1] of BASELINE.json, GPT-2-shaped
byte-level BPE, 1 GiB of synthetic UTF-8 per GPU).  One process per GPU, documents sharded by rank with NO collective on
the data path (weak scaling: every rank encodes its own shard); NCCL is used only for the barrier and the max-over-ranks
of the timings.  `value` is measured with inputs resident in HBM (CUDA events on the stream the kernels run on);
`e2e` goes through the host-buffer C-ABI call (tkz_encode_batch: pinned H2D of the text, D2H of the encoding).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (tokenizer, corpus, description, truncation, padding, docs per sub-batch (0 = whole shard))
    "c2b": ("gpt2_whitespace", "c2", "GPT-2-shaped byte-level BPE 50,257/50k merges, pre_tokenizer Whitespace (words split on whitespace)", None, None, 0),
    "c2a": ("gpt2_bytelevel", "c2", "GPT-2-shaped byte-level BPE 50,257/50k merges, pre_tokenizer ByteLevel (reference: null => whole document = one pre-token)", None, None, 0),
    "c3": ("bert_wordpiece", "c3", "BERT-shaped WordPiece 30,522, ASCII-lowercase + ws/punct split, truncate/pad 512", 512, {"length": 512, "pad_id": 0}, 2097152),
    "c4b": ("llama3_whitespace", "c4", "Llama-3-shaped byte-level BPE 128,256, pre_tokenizer Whitespace, multilingual", None, None, 0),
    "c4a": ("llama3_sequence", "c4", "Llama-3-shaped byte-level BPE 128,256, pre_tokenizer Sequence (reference: null => whole document)", None, None, 0),
    "c5b": ("gpt2_whitespace", "c5", "GPT-2-shaped BPE, skewed documents 1 B..4 MiB with long unbroken words, Whitespace", None, None, 0),
    "c5a": ("gpt2_bytelevel", "c5", "GPT-2-shaped BPE, skewed documents 1 B..4 MiB, whole-document pre-tokens", None, None, 0),
}
DEFAULT_MIB = {"c2b": 1024, "c2a": 1024, "c3": 1024, "c4b": 2048, "c4a": 2048, "c5b": 1024, "c5a": 256}


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        while not self.stop_flag.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([x.strip() for x in line.split(",")])
        p.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def make_corpus(cname, nbytes, seed, pinned=True):
    from tools import corpus
    out = None
    if pinned:
        import torch
        out = torch.empty(nbytes + 4096, dtype=torch.uint8, pin_memory=True).numpy()
    text, off = corpus.generate(cname, nbytes, seed, out=out)
    return text, off


def sub_batches(off, docs_per_batch):
    nd = len(off) - 1
    if docs_per_batch <= 0 or docs_per_batch >= nd:
        return [(0, nd)]
    return [(a, min(nd, a + docs_per_batch)) for a in range(0, nd, docs_per_batch)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(tok_name, text, off, trunc, pad, algo, budget_s=15.0):
    """The reference algorithm (C restatement, oracle/) on the host cores, bounded sample of the same workload."""
    from oracle import oracle as orc
    from tools import tokenizers_io
    o = orc.OracleTokenizer.from_json(tokenizers_io.tokenizer_json(tok_name))
    o.truncation = trunc
    o.padding = pad
    cores = os.cpu_count() or 1
    nd = len(off) - 1
    # calibrate on a small prefix, then size the sample for ~budget_s of wall clock
    probe_docs = max(1, min(nd, int(np.searchsorted(off, min(int(off[-1]), 256 << 10)))))
    probe_docs = max(probe_docs, min(nd, cores * 4))
    t0 = time.perf_counter()
    o.count_tokens(text, off[: probe_docs + 1], algo=algo, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = float(off[probe_docs]) / dt
    want = int(min(float(off[-1]), rate * budget_s))
    ndocs = max(probe_docs, min(nd, int(np.searchsorted(off, want))))
    t0 = time.perf_counter()
    ntok = o.count_tokens(text, off[: ndocs + 1], algo=algo, threads=cores)
    dt = time.perf_counter() - t0
    nbytes = int(off[ndocs])
    return {"value": nbytes / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"first {ndocs} documents ({nbytes / 2**20:.1f} MiB) of the rank-0 shard, {dt:.1f} s, algo={'BPE.tokenize literal' if algo == 0 else 'fast-exact'}",
            "tokens_per_s": ntok / dt, "note": "C restatement of tokenizer-zig's Tokenizer.encode (oracle/), not the Zig binary: no zig toolchain in the image"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    tok_name, cname, desc, trunc, pad, _ = WORKLOADS[args.workload]
    size = args.size_mib << 20
    text, off = make_corpus(cname, min(size, 256 << 20), 1234, pinned=False)
    from oracle import oracle as orc
    from tools import tokenizers_io
    o = orc.OracleTokenizer.from_json(tokenizers_io.tokenizer_json(tok_name))
    o.truncation = trunc
    o.padding = pad
    cores = os.cpu_count() or 1
    algo = 0
    nd = len(off) - 1
    # size one step for ~8 s of wall clock
    probe = max(1, min(nd, cores * 8))
    t0 = time.perf_counter(); o.count_tokens(text, off[: probe + 1], algo=algo, threads=cores); dt = max(time.perf_counter() - t0, 1e-4)
    rate = float(off[probe]) / dt
    ndocs = max(probe, min(nd, int(np.searchsorted(off, min(float(off[-1]), rate * args.ref_step_s)))))
    sub = off[: ndocs + 1]
    nbytes = int(sub[-1])
    for _ in range(args.warmup):
        o.count_tokens(text, sub, algo=algo, threads=cores)
    t0 = time.perf_counter()
    ntok = 0
    for _ in range(args.steps):
        ntok = o.count_tokens(text, sub, algo=algo, threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = nbytes / dt / 1e9
    line = {"impl": "reference", "metric": "encode_input_throughput", "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "tokens_per_s": ntok / dt,
            "config": {"workload": f"{args.workload}: {desc}", "tokenizer": tok_name, "corpus": cname, "step_sample_bytes": nbytes, "step_sample_docs": ndocs},
            "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": f"each step = first {ndocs} documents ({nbytes / 2**20:.1f} MiB) of the seed-1234 shard, all {cores} host threads, BPE.tokenize/WordPiece.tokenize literal restatement"},
            "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import tokzig_b200 as tz
    from tools import tokenizers_io

    tok_name, cname, desc, trunc, pad, docs_per_batch = WORKLOADS[args.workload]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    size = args.size_mib << 20
    t_gen = time.time()
    text, off = make_corpus(cname, size, 1234 + rank)
    t_gen = time.time() - t_gen
    nd = len(off) - 1
    nbytes = int(off[-1])

    stream = torch.cuda.current_stream()
    tok = tz.Tokenizer.from_json(tokenizers_io.tokenizer_json(tok_name), device=local_rank, stream=stream.cuda_stream)
    L = tz.lib()
    ctx_h = tok.context_handle()
    params = tz.EncodeParams()
    if trunc is not None:
        params.has_truncation, params.max_length = 1, trunc
    if pad is not None:
        params.has_padding, params.pad_length, params.pad_id = 1, pad["length"], pad.get("pad_id", 0)
    params.outputs = args.outputs
    tok.truncation = None if trunc is None else {"max_length": trunc}
    tok.padding = pad

    d_text = torch.from_numpy(text).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    batches = sub_batches(off, docs_per_batch)
    # per sub-batch rebased doc offsets on the device
    d_offs = []
    for a, b in batches:
        d_offs.append((d_off[a: b + 1] - d_off[a]).contiguous())
    import ctypes as C

    stats = tz.Stats()
    agg = {"tokens": 0, "real": 0, "launches": 0, "ms": [0.0] * 5, "words": 0}

    def device_step(collect=False):
        tot_tokens = tot_real = launches = words = 0
        ms = [0.0] * 5
        for (a, b), doff in zip(batches, d_offs):
            r = tz.BatchResult()
            base = int(off[a]); nb = int(off[b]) - base
            rc = L.tkz_encode_batch_device(ctx_h, C.c_void_p(d_text.data_ptr() + base), C.c_void_p(doff.data_ptr()), b - a, nb, C.byref(params), C.byref(r))
            if rc != 0:
                raise RuntimeError(f"tkz_encode_batch_device rc={rc}: {L.tkz_last_error(ctx_h)}")
            tot_tokens += r.n_tokens; tot_real += r.n_real_tokens
            if collect:
                L.tkz_ctx_get_stats(ctx_h, C.byref(stats))
                launches += stats.kernel_launches; words += stats.n_words
                agg["uniq"] = agg.get("uniq", 0) * 0 + int(stats.n_unique_words); agg["long"] = int(stats.n_long_words)
                agg["path"] = int(stats.path)
                for i, k in enumerate(("ms_split", "ms_model", "ms_scan", "ms_emit", "ms_total")):
                    ms[i] += getattr(stats, k)
        return tot_tokens, tot_real, launches, ms, words

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- parity spot-check against the oracle (untimed, rank 0): the first documents of the shard
    parity = None
    if rank == 0 and not args.no_verify:
        from oracle import oracle as orc
        o = orc.OracleTokenizer.from_json(tokenizers_io.tokenizer_json(tok_name))
        o.truncation = trunc; o.padding = pad
        k = max(1, min(nd, int(np.searchsorted(off, 2 << 20))))
        ref = o.encode_packed(text[: int(off[k])], off[: k + 1], algo=1, threads=os.cpu_count() or 1)
        got = tok.encode_packed(text[: int(off[k])], off[: k + 1])
        parity = bool(np.array_equal(got.ids, ref.ids) and np.array_equal(got.offsets, ref.offsets) and np.array_equal(got.doc_tok_off, ref.doc_tok_off)
                      and np.array_equal(got.attention_mask, ref.attention_mask))
        if not parity:
            raise SystemExit("PARITY FAILURE: GPU encoding differs from the oracle on the verification sample")

    # ---- device-resident timing
    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for s in range(args.steps):
        tot_tokens, tot_real, launches, ms, words = device_step(collect=True)
        agg["tokens"] = tot_tokens; agg["real"] = tot_real; agg["launches"] += launches; agg["words"] = words
        agg["ms"] = [x + y for x, y in zip(agg["ms"], ms)]
    ev1.record(stream)
    torch.cuda.synchronize(); barrier()
    sampler.stop_flag.set()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(nbytes), float(agg["real"]), float(agg["tokens"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step = float(t.item()) / args.steps
    all_bytes, all_real, all_slots = (float(x) for x in tot.tolist())

    # ---- end to end through the host-buffer C-ABI call (pinned text in, encoding out to host)
    e2e = None
    if not args.no_e2e:
        off_batches = [(off[a: b + 1] - off[a]).astype(np.uint64) for a, b in batches]

        def host_step():
            h2d = d2h = 0
            for (a, b), ob in zip(batches, off_batches):
                r = tz.BatchResult()
                base = int(off[a])
                rc = L.tkz_encode_batch(ctx_h, C.c_void_p(text.ctypes.data + base), C.c_void_p(ob.ctypes.data), b - a, C.byref(params), C.byref(r))
                if rc != 0:
                    raise RuntimeError(f"tkz_encode_batch rc={rc}: {L.tkz_last_error(ctx_h)}")
                h2d += int(ob[-1]) + ob.nbytes
                per_slot = 4 + (8 if params.outputs & 2 else 0) + (4 if params.outputs & 4 else 0) + (4 if params.outputs & 8 else 0) + (4 if params.outputs & 16 else 0)
                d2h += int(r.n_tokens) * per_slot + (b - a + 1) * 8
            return h2d, d2h

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        host_step()                       # warm-up: sizes the pinned result buffers
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h2d, d2h = host_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e = {"value": all_bytes / float(td.item()) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": float(td.item()) * 1e3, "steps": e2e_steps, "tokens_per_s": all_real / float(td.item()),
               "api": "tkz_encode_batch (host pointers; pinned text H2D + result D2H inside the timed region)"}
        if params.outputs != 1 and pad is None:
            # the same call asking for the ids only (what the metric's "output ids" names): shows how much of e2e is the D2H of
            # offsets + attention mask
            full_mask = params.outputs
            params.outputs = 1
            host_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                h2d1, d2h1 = host_step()
            torch.cuda.synchronize()
            dt1 = (time.perf_counter() - t0) / 2
            params.outputs = full_mask
            e2e["ids_only"] = {"value": nbytes / dt1 / 1e9, "unit": "GB/s", "ms_per_step": dt1 * 1e3, "d2h_bytes_per_step": d2h1, "note": "rank-local"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # roofline of the dominant kernel stage: algorithmic bytes of one step / that stage's device time in one step
    per_step_ms = [x / args.steps for x in agg["ms"]]
    names = ["split (K0+K1)", "model (K3 bpe | K4 wordpiece)", "scan", "emit (K5)"]
    path_name = {0: "per-occurrence pipeline", 1: "dedup multi-pass pipeline", 2: "slice pipeline (2 passes)"}.get(agg.get("path"), "?")
    if agg.get("path") == 2:
        names[0] = "slice_words_kernel (pass A: normalise + split + word-table probe + inline model)"
        names[1] = "word-list kernels (pre-tokens > 256 B)"
        names[3] = "slice_emit_kernel (pass B: fromTokens + truncate + pad)"
    dom = int(np.argmax(per_step_ms[:4]))
    b_alg = nbytes + 4 * agg["tokens"]                       # input bytes + 4 B x id slots written (one rank)
    b_full = nbytes + (4 + (8 if params.outputs & 2 else 0) + (4 if params.outputs & 4 else 0)) * agg["tokens"]
    ach_dom = b_alg / (per_step_ms[dom] * 1e-3) / 1e9 if per_step_ms[dom] > 0 else 0.0
    ach_pipe = b_alg / (ms_step * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:{args.size_mib}")
        if tj and agg.get("path") == 2:
            traffic = tj.get("slice_words_kernel" if dom == 0 else ("slice_emit_kernel" if dom == 3 else ""))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": ach_dom, "peak": peak, "unit": "GB/s", "frac": ach_dom / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_step": b_alg,
                "pipeline": {"achieved": ach_pipe, "frac": ach_pipe / peak, "note": "all kernels of a step: input bytes + 4 B per id slot over the whole device time"},
                "stage_ms_per_step": dict(zip(["split", "model", "scan", "emit", "total_kernels"], [round(x, 4) for x in per_step_ms])),
                "full_output_bytes_per_step": b_full}
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(tok_name, text, off, trunc, pad, algo=0)
    line = {"metric": "encode_input_throughput", "value": all_bytes / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": f"synthetic ({cname} generator, seed 1234+rank, generated in {t_gen:.1f} s)",
            "tokens_per_s": all_real / (ms_step * 1e-3), "slots_per_s": all_slots / (ms_step * 1e-3),
            "config": {"workload": f"{args.workload}: {desc}", "tokenizer": tok_name, "corpus": cname, "bytes_per_gpu": nbytes, "docs_per_gpu": nd,
                       "words_per_gpu": agg["words"], "unique_words_last_batch": agg.get("uniq"), "long_words_last_batch": agg.get("long"), "tokens_per_gpu": agg["real"], "sub_batches": len(batches), "outputs_mask": int(params.outputs), "pipeline": path_name,
                       "l2": "inputs (>= 1 GiB per step) larger than the 126 MB L2; no flush needed", "parallelism": f"documents sharded x{world}, no collective"},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(agg["launches"]), "clocks": sampler.summary(),
            "parity_checked_vs_oracle": parity}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2b", choices=sorted(WORKLOADS))
    ap.add_argument("--size-mib", type=int, default=0)
    ap.add_argument("--outputs", type=int, default=7, help="TKZ_OUT_* mask: ids|offsets|attention = 7 (the Encoding the north star names)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-step-s", type=float, default=8.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    if args.size_mib <= 0:
        args.size_mib = DEFAULT_MIB[args.workload]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world > 1 and rank != 0:
            time.sleep(0.5)
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
