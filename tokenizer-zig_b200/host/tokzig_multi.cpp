// tokzig_multi.cpp -- the multi-GPU entry point of the batch encoder (tkzm_*): N contexts on the N GPUs of one box, one host
// thread per GPU, documents cut into contiguous shards with NO collective on the data path.
//
// The reference has no counterpart with data on several devices; its closest analogue is the per-thread arena pool
// (src/arena.zig:252-335: one TokenizerArena per worker thread, models shared read-only).  Here the model tables are
// replicated per GPU, every context owns its device arenas and its pinned result buffers, and the host side only decides
// WHERE to cut: by bytes (north star: "byte-balanced ranges"), or by a modelled cost when the bytes of a corpus are not
// equally expensive (BASELINE config 5: documents with MiB-long unbroken words).
//
// Nothing here tokenizes: encode is N concurrent tkz_encode_batch_compact calls.
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tokzig_b200.h"

struct tkzm_pool {
    std::vector<tkz_ctx*> ctx;
    std::vector<int> device;
    std::string err;
    bool has_pretok = false;
    uint8_t raw_class[256];                 // class of the RAW byte under the uploaded model (TKZ_CLS_*), for the cost model
    std::vector<std::vector<uint64_t>> off;  // per shard: doc offsets rebased to the shard's first byte
};

namespace {

// modelled device time per input byte, relative to ordinary text (B200: ordinary text ~5 ms / GiB; pre-tokens of 65 B .. 12 KiB
// in the windowed block kernels ~100 ms / GiB; longer ones on the cooperative grid ~43 ms per 245 MB)
constexpr double COST_LONG_WORD = 20.0, COST_HUGE_WORD = 36.0;
constexpr uint64_t LONG_MIN = 65, HUGE_MIN = 12289;

// best effort: run the calling thread on the CPUs of the GPU's NUMA node, so that the pinned buffers its context allocates
// (first touch) and the staging copies stay local to the GPU's PCIe root
void bind_to_gpu_node(tkz_ctx* c) {
    const int node = tkz_ctx_numa_node(c);
    if (node < 0) return;
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return;
    char buf[4096];
    const size_t n = fread(buf, 1, sizeof buf - 1, f);
    fclose(f);
    buf[n] = 0;
    cpu_set_t set; CPU_ZERO(&set);
    int any = 0;
    for (char* p = buf; *p;) {
        char* e; long a = strtol(p, &e, 10); if (e == p) break;
        long b = a;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long c2 = a; c2 <= b && c2 < CPU_SETSIZE; c2++) { CPU_SET((int)c2, &set); any = 1; }
        p = (*e == ',') ? e + 1 : e;
        if (*e != ',' ) break;
    }
    if (any) sched_setaffinity(0, sizeof set, &set);
}

}  // namespace

extern "C" int tkzm_create(const int32_t* devices, int32_t n, tkzm_pool** out) {
    if (!out || !devices || n <= 0) return TKZ_ERR_INVALID_ARG;
    *out = nullptr;
    tkzm_pool* p = new tkzm_pool();
    memset(p->raw_class, 0, sizeof p->raw_class);
    for (int i = 0; i < n; i++) {
        tkz_ctx* c = nullptr;
        const int rc = tkz_ctx_create(devices[i], nullptr, 0, &c);
        if (rc != TKZ_OK) {
            for (tkz_ctx* x : p->ctx) tkz_ctx_destroy(x);
            delete p;
            return rc;
        }
        p->ctx.push_back(c); p->device.push_back(devices[i]);
    }
    p->off.resize(n);
    *out = p;
    return TKZ_OK;
}

extern "C" void tkzm_destroy(tkzm_pool* p) {
    if (!p) return;
    for (tkz_ctx* c : p->ctx) tkz_ctx_destroy(c);
    delete p;
}

extern "C" int32_t tkzm_size(tkzm_pool* p) { return p ? (int32_t)p->ctx.size() : 0; }
extern "C" tkz_ctx* tkzm_ctx(tkzm_pool* p, int32_t i) { return (p && i >= 0 && i < (int32_t)p->ctx.size()) ? p->ctx[i] : nullptr; }
extern "C" const char* tkzm_last_error(tkzm_pool* p) { return p ? p->err.c_str() : "no pool"; }

extern "C" int tkzm_model_upload(tkzm_pool* p, const tkz_model_desc* d) {
    if (!p || !d) return TKZ_ERR_INVALID_ARG;
    std::vector<int> rcs(p->ctx.size(), TKZ_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < p->ctx.size(); i++) th.emplace_back([&, i] { rcs[i] = tkz_model_upload(p->ctx[i], d); });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < rcs.size(); i++) if (rcs[i] != TKZ_OK) { p->err = std::string("GPU ") + std::to_string(p->device[i]) + ": " + tkz_last_error(p->ctx[i]); return rcs[i]; }
    p->has_pretok = d->class_lut != nullptr;
    for (int b = 0; b < 256; b++) {
        const uint16_t nb = d->norm_lut ? d->norm_lut[b] : (uint16_t)b;
        p->raw_class[b] = nb == TKZ_NORM_DROP ? (uint8_t)TKZ_CLS_DELIM : (d->class_lut ? d->class_lut[nb & 0xFF] : (uint8_t)TKZ_CLS_WORD);
    }
    return TKZ_OK;
}

// shard k takes documents [bounds[k], bounds[k+1]): cut at the document boundary nearest to k * total / n_shards of the
// measure (bytes, or `cost` per document)
extern "C" int tkzm_shard_bounds(const uint64_t* doc_off, uint64_t n_docs, int32_t n_shards, const double* cost, uint64_t* bounds) {
    if (!doc_off || !bounds || n_shards <= 0) return TKZ_ERR_INVALID_ARG;
    std::vector<double> acc;
    if (cost) { acc.resize(n_docs + 1); acc[0] = 0.0; for (uint64_t d = 0; d < n_docs; d++) acc[d + 1] = acc[d] + cost[d]; }
    auto at = [&](uint64_t i) -> double { return cost ? acc[i] : (double)(doc_off[i] - doc_off[0]); };
    const double total = at(n_docs);
    bounds[0] = 0;
    for (int32_t k = 1; k < n_shards; k++) {
        const double target = total * (double)k / (double)n_shards;
        uint64_t lo = 0, hi = n_docs + 1;                       // first i with at(i) >= target
        while (lo < hi) { const uint64_t mid = lo + (hi - lo) / 2; if (at(mid) < target) lo = mid + 1; else hi = mid; }
        uint64_t i = lo;
        if (i > 0 && i <= n_docs && (target - at(i - 1)) < (at(std::min(i, n_docs)) - target)) i--;
        bounds[k] = std::max(bounds[k - 1], std::min(i, n_docs));
    }
    bounds[n_shards] = n_docs;
    return TKZ_OK;
}

// cost[d] = bytes of the document + (COST_LONG_WORD - 1) x its bytes in pre-tokens of 65 .. 12288 bytes + (COST_HUGE_WORD - 1)
// x its bytes in longer ones.  raw_class = TKZ_CLS_* of every raw byte, NULL = no pre-tokenizer (a document is one pre-token).
// One pass over the text on `threads` host threads (0 = all).
extern "C" int tkzm_document_costs(const uint8_t* text, const uint64_t* doc_off, uint64_t n_docs, const uint8_t* raw_class, int32_t threads, double* cost) {
    if (!doc_off || !cost || (n_docs && doc_off[n_docs] > doc_off[0] && !text)) return TKZ_ERR_INVALID_ARG;
    if (threads <= 0) threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    auto extra = [](uint64_t wlen) -> double {
        return wlen >= HUGE_MIN ? (COST_HUGE_WORD - 1.0) * (double)wlen : (wlen >= LONG_MIN ? (COST_LONG_WORD - 1.0) * (double)wlen : 0.0);
    };
    auto work = [&](uint64_t d0, uint64_t d1) {
        for (uint64_t d = d0; d < d1; d++) {
            const uint64_t a = doc_off[d], b = doc_off[d + 1];
            double c = (double)(b - a);
            if (!raw_class) c += extra(b - a);
            else if (b - a >= LONG_MIN) {
                // only runs of LONG_MIN or more WORD bytes matter: every such run holds a byte at a multiple of LONG_MIN from the
                // document start, so only those bytes are looked at, and a WORD byte there is extended to its run
                uint64_t i = a + LONG_MIN - 1;
                while (i < b) {
                    if (raw_class[text[i]] != TKZ_CLS_WORD) { i += LONG_MIN; continue; }
                    uint64_t lo = i, hi = i + 1;
                    while (lo > a && raw_class[text[lo - 1]] == TKZ_CLS_WORD) lo--;
                    while (hi < b && raw_class[text[hi]] == TKZ_CLS_WORD) hi++;
                    if (hi - lo >= LONG_MIN) c += extra(hi - lo);
                    i = hi + LONG_MIN;                             // the next run starts behind hi
                }
            }
            cost[d] = c;
        }
    };
    if (threads == 1 || n_docs < 1024) { work(0, n_docs); return TKZ_OK; }
    // byte-balanced document ranges per thread
    std::vector<uint64_t> b((size_t)threads + 1);
    tkzm_shard_bounds(doc_off, n_docs, threads, nullptr, b.data());
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(work, b[t], b[t + 1]);
    for (auto& t : th) t.join();
    return TKZ_OK;
}

extern "C" int tkzm_encode_batch_compact(tkzm_pool* p, const uint8_t* text, const uint64_t* doc_off, uint64_t n_docs, const tkz_encode_params* params,
                                         int want_offsets, int cost_balanced, uint64_t* bounds, tkz_compact_result* results, double* shard_ms) {
    if (!p || !doc_off || !bounds || !results) return TKZ_ERR_INVALID_ARG;
    const int32_t n = (int32_t)p->ctx.size();
    if (cost_balanced == 2) {
        // the caller's cut
        if (bounds[0] != 0 || bounds[n] != n_docs) { p->err = "bounds must run from 0 to n_docs"; return TKZ_ERR_INVALID_ARG; }
        for (int32_t k = 0; k < n; k++) if (bounds[k] > bounds[k + 1]) { p->err = "bounds must not decrease"; return TKZ_ERR_INVALID_ARG; }
    } else if (cost_balanced) {
        std::vector<double> cost(n_docs);
        int rc = tkzm_document_costs(text, doc_off, n_docs, p->has_pretok ? p->raw_class : nullptr, 0, cost.data());
        if (rc == TKZ_OK) rc = tkzm_shard_bounds(doc_off, n_docs, n, cost.data(), bounds);
        if (rc != TKZ_OK) return rc;
    } else {
        const int rc = tkzm_shard_bounds(doc_off, n_docs, n, nullptr, bounds);
        if (rc != TKZ_OK) return rc;
    }
    std::vector<int> rcs(n, TKZ_OK);
    std::vector<std::thread> th;
    for (int32_t k = 0; k < n; k++) {
        th.emplace_back([&, k] {
            bind_to_gpu_node(p->ctx[k]);
            const uint64_t d0 = bounds[k], d1 = bounds[k + 1], base = doc_off[d0];
            std::vector<uint64_t>& off = p->off[k];
            off.resize(d1 - d0 + 1);
            for (uint64_t i = 0; i <= d1 - d0; i++) off[i] = doc_off[d0 + i] - base;
            const auto t0 = std::chrono::steady_clock::now();
            rcs[k] = tkz_encode_batch_compact(p->ctx[k], text ? text + base : nullptr, off.data(), d1 - d0, params, want_offsets, &results[k]);
            if (results[k].err_doc >= 0) results[k].err_doc += (int64_t)d0;
            if (shard_ms) shard_ms[k] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        });
    }
    for (auto& t : th) t.join();
    // the first failing shard in document order decides the error (as a loop over Tokenizer.encode would)
    for (int32_t k = 0; k < n; k++) if (rcs[k] != TKZ_OK) { p->err = std::string("shard ") + std::to_string(k) + " (GPU " + std::to_string(p->device[k]) + "): " + tkz_last_error(p->ctx[k]); return rcs[k]; }
    return TKZ_OK;
}
