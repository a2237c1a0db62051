//! gpu_tokenizer.zig -- the part of src/lib.zig that changes: `Tokenizer.encode` (src/lib.zig:109-160) becomes one
//! C-ABI batch call.  Everything else of the library (config.zig loader, Vocab, decode, lookups) stays as it is.
//!
//! Drop-in: `GpuTokenizer.fromJson` builds the unchanged `Tokenizer` (so tokenToId / idToToken / decode / getVocabSize
//! keep working through `.base`), flattens its model into `cuda.ModelDesc`, uploads it once, and `encode` /
//! `encodeBatch` return the same owned `Encoding` the reference returns (src/encoding.zig:231-243).
//! NOT COMPILED in the build image (no zig toolchain); logic lives behind the C ABI, this file only marshals.

const std = @import("std");
const lib = @import("lib.zig");
const cuda = @import("cuda.zig");
const bpe = @import("model/bpe.zig");
const wordpiece = @import("model/wordpiece.zig");

pub const GpuError = error{ OutOfMemory, MissingUnkToken, InvalidUtf8, CudaError, InvalidArgument };

fn check(rc: c_int) GpuError!void {
    return switch (rc) {
        cuda.OK => {},
        cuda.ERR_OOM => error.OutOfMemory,
        cuda.ERR_MISSING_UNK => error.MissingUnkToken,
        cuda.ERR_INVALID_UTF8 => error.InvalidUtf8,
        cuda.ERR_INVALID_ARG => error.InvalidArgument,
        else => error.CudaError,
    };
}

pub const GpuTokenizer = struct {
    base: lib.Tokenizer,
    ctx: *cuda.Ctx,

    const Self = @This();

    /// Same signature family as FastTokenizer.fromJson (src/lib.zig:275): the JSON is parsed a second time to learn the
    /// component types, exactly as FastTokenizer does for the model type (src/lib.zig:281-293), because the loader's
    /// normalizer / pre-tokenizer vtables carry no type tag (src/config.zig:347-357, 388-398).
    pub fn fromJson(allocator: std.mem.Allocator, json_content: []const u8, device: c_int) !Self {
        var base = try lib.Tokenizer.fromJson(allocator, json_content);
        errdefer base.deinit();

        const parsed = std.json.parseFromSlice(std.json.Value, allocator, json_content, .{}) catch return error.InvalidJson;
        defer parsed.deinit();
        const root = parsed.value.object;
        const model_obj = (root.get("model") orelse return error.MissingModel).object;
        const is_bpe = if (model_obj.get("type")) |t| std.mem.eql(u8, t.string, "BPE") else false;

        // byte map and class table (the reference's components are all byte-wise, SURVEY.md section 0)
        var norm_lut: [256]u16 = undefined;
        var class_lut: [256]u8 = undefined;
        var has_norm = false;
        var has_pretok = false;
        if (componentType(root, "normalizer")) |t| {
            if (std.mem.eql(u8, t, "BertNormalizer") or std.mem.eql(u8, t, "Lowercase")) has_norm = true; // config.zig:345-358
        }
        var bert_split = false;
        if (componentType(root, "pre_tokenizer")) |t| {
            if (std.mem.eql(u8, t, "BertPreTokenizer")) {
                has_pretok = true;
                bert_split = true;
            } else if (std.mem.eql(u8, t, "Whitespace") or std.mem.eql(u8, t, "WhitespaceSplit")) has_pretok = true; // config.zig:387-399
        }
        for (0..256) |i| {
            const b: u8 = @intCast(i);
            norm_lut[i] = std.ascii.toLower(b); // config.zig:364-379
            class_lut[i] = if (bert_split)
                (if (std.ascii.isWhitespace(b)) cuda.CLS_DELIM else if (isPunct(b)) cuda.CLS_ISOLATE else cuda.CLS_WORD) // config.zig:405-438
            else
                (if (b == ' ' or b == '\t' or b == '\n' or b == '\r') cuda.CLS_DELIM else cuda.CLS_WORD); // config.zig:440-450
        }

        // flatten the model (src/model/bpe.zig:36-46, src/model/wordpiece.zig:13-20)
        var keys = std.ArrayListUnmanaged(u8){};
        defer keys.deinit(allocator);
        var offs = std.ArrayListUnmanaged(u64){};
        defer offs.deinit(allocator);
        var ids = std.ArrayListUnmanaged(u32){};
        defer ids.deinit(allocator);
        var mf = std.ArrayListUnmanaged(u32){};
        defer mf.deinit(allocator);
        var ms = std.ArrayListUnmanaged(u32){};
        defer ms.deinit(allocator);
        var mr = std.ArrayListUnmanaged(u32){};
        defer mr.deinit(allocator);
        var mn = std.ArrayListUnmanaged(u32){};
        defer mn.deinit(allocator);
        try offs.append(allocator, 0);

        var desc = std.mem.zeroes(cuda.ModelDesc);
        if (is_bpe) {
            const m: *bpe.BPE = @ptrCast(@alignCast(base.model_ptr.?));
            var it = m.vocab.iterator();
            while (it.next()) |e| {
                try keys.appendSlice(allocator, e.key_ptr.*);
                try offs.append(allocator, keys.items.len);
                try ids.append(allocator, e.value_ptr.*);
            }
            var mit = m.merges.iterator();
            while (mit.next()) |e| {
                try mf.append(allocator, @intCast(e.key_ptr.* >> 32));
                try ms.append(allocator, @truncate(e.key_ptr.*));
                try mr.append(allocator, e.value_ptr.rank);
                try mn.append(allocator, e.value_ptr.new_id);
            }
            desc.model_kind = cuda.MODEL_BPE;
            if (m.unk_token) |u| {
                if (m.vocab.get(u)) |uid| { // bpe.zig:198-207
                    desc.has_unk = 1;
                    desc.unk_id = uid;
                }
            }
        } else {
            const m: *wordpiece.WordPiece = @ptrCast(@alignCast(base.model_ptr.?));
            var it = m.vocab.iterator();
            while (it.next()) |e| {
                try keys.appendSlice(allocator, e.key_ptr.*);
                try offs.append(allocator, keys.items.len);
                try ids.append(allocator, e.value_ptr.*);
            }
            desc.model_kind = cuda.MODEL_WORDPIECE;
            if (m.vocab.get(m.unk_token)) |uid| { // wordpiece.zig:150,212
                desc.has_unk = 1;
                desc.unk_id = uid;
            }
            desc.prefix = m.continuing_subword_prefix.ptr;
            desc.prefix_len = @intCast(m.continuing_subword_prefix.len);
            desc.max_input_chars_per_word = m.max_input_chars_per_word;
        }
        desc.norm_lut = if (has_norm) &norm_lut else null;
        desc.class_lut = if (has_pretok) &class_lut else null;
        desc.vocab_bytes = keys.items.ptr;
        desc.vocab_off = offs.items.ptr;
        desc.vocab_ids = ids.items.ptr;
        desc.vocab_n = @intCast(ids.items.len);
        desc.merge_first = mf.items.ptr;
        desc.merge_second = ms.items.ptr;
        desc.merge_rank = mr.items.ptr;
        desc.merge_new = mn.items.ptr;
        desc.merges_n = @intCast(mf.items.len);

        var ctx: ?*cuda.Ctx = null;
        try check(cuda.tkz_ctx_create(device, null, 0, &ctx));
        errdefer cuda.tkz_ctx_destroy(ctx);
        try check(cuda.tkz_model_upload(ctx.?, &desc));
        return .{ .base = base, .ctx = ctx.? };
    }

    pub fn deinit(self: *Self) void {
        cuda.tkz_ctx_destroy(self.ctx);
        self.base.deinit();
    }

    fn params(self: *const Self) cuda.EncodeParams {
        var p = cuda.EncodeParams{};
        if (self.base.truncation) |t| { // lib.zig:150-152
            p.has_truncation = 1;
            p.max_length = t.max_length;
        }
        if (self.base.padding) |pd| { // lib.zig:155-157, encoding.zig:386
            if (pd.length) |len| {
                p.has_padding = 1;
                p.pad_length = len;
                p.pad_id = pd.pad_id;
                p.pad_type_id = pd.pad_type_id;
                p.pad_left = if (pd.direction == .left) 1 else 0;
            }
        }
        return p;
    }

    /// Tokenizer.encode (src/lib.zig:109): same result type, owned by the caller (`encoding.deinit()`).
    pub fn encode(self: *Self, text: []const u8, add_special_tokens: bool) !lib.Encoding {
        const one = [_][]const u8{text};
        const encs = try self.encodeBatch(&one, add_special_tokens);
        defer self.base.allocator.free(encs);
        return encs[0];
    }

    /// Replaces the byte tables derived from the JSON component types: for callers that hand-wire `normalizer_impl` /
    /// `pretokenizer_impl` (public fields, src/lib.zig:37-38) with the struct variants of src/normalizer/normalizer.zig and
    /// src/pretokenizer/pretokenizer.zig.  Every such component is byte-wise, so any chain composes into one byte map
    /// (norm_lut[b] = normalised byte or cuda.NORM_DROP) and one class table over normalised bytes; null = component absent.
    /// (The C++ host mirror composes these tables itself, tkzh_set_normalizer / tkzh_set_pretokenizer; this Zig layer reads
    /// only the JSON types, see INTEGRATION.md "Known divergences".)
    pub fn setByteTables(self: *Self, desc_base: *cuda.ModelDesc, norm_lut: ?*const [256]u16, class_lut: ?*const [256]u8) !void {
        desc_base.norm_lut = if (norm_lut) |p| p else null;
        desc_base.class_lut = if (class_lut) |p| p else null;
        try check(cuda.tkz_model_upload(self.ctx, desc_base));
    }

    /// Borrowed view of a batch encoding: the arrays belong to the context's pinned buffers and stay valid until the next
    /// encode on this tokenizer -- the contract of FastTokenizer.encode / SpanEncoding (src/lib.zig:353-356).  Nothing is
    /// allocated per document; masks and padding slots are not stored at all (they are constants of `kept` and the padding
    /// parameters) and come out of `expandInto`.
    pub const BatchView = struct {
        r: cuda.CompactResult,

        pub fn len(self: *const BatchView) usize {
            return @intCast(self.r.n_docs);
        }
        /// kept real tokens of document d (after truncation, before padding)
        pub fn kept(self: *const BatchView, d: usize) usize {
            return @intCast(self.r.doc_kept_off.?[d + 1] - self.r.doc_kept_off.?[d]);
        }
        pub fn id(self: *const BatchView, d: usize, i: usize) u32 {
            const k: usize = @intCast(self.r.doc_kept_off.?[d]);
            return if (self.r.ids16) |p| p[k + i] else self.r.ids.?[k + i];
        }
        pub fn offset(self: *const BatchView, d: usize, i: usize) lib.Offset {
            const k: usize = @intCast(self.r.doc_kept_off.?[d]);
            if (self.r.offsets_packed) |p| {
                const v = p[k + i];
                if (v != 0xFFFF) return lib.Offset.init(v & 0xFF, v >> 8);
                // a token of a pre-token of 256+ bytes: binary search in the side list (sorted by kept index)
                const w = self.r.wide_tokens.?;
                var lo: usize = 0;
                var hi: usize = @intCast(self.r.n_wide);
                const slot: u64 = k + i;
                while (lo < hi) {
                    const mid = (lo + hi) / 2;
                    const sv = (@as(u64, w[4 * mid + 1]) << 32) | w[4 * mid];
                    if (sv < slot) lo = mid + 1 else hi = mid;
                }
                return lib.Offset.init(w[4 * lo + 2], w[4 * lo + 3]);
            }
            return lib.Offset.init(self.r.offsets.?[2 * (k + i)], self.r.offsets.?[2 * (k + i) + 1]);
        }
        /// slots of document d after Encoding.pad (src/encoding.zig:385-463)
        pub fn slots(self: *const BatchView, d: usize) usize {
            return @intCast(cuda.tkz_compact_slots(&self.r, d, d + 1));
        }
        /// Encoding.fromTokens + pad of document d into caller-owned arrays of `slots(d)` entries (offsets: 2 u32 per slot)
        pub fn expandInto(self: *const BatchView, d: usize, ids: []u32, offsets: []u32, attention: []u32, type_ids: []u32, special: []u32) !void {
            try check(cuda.tkz_compact_expand(&self.r, d, d + 1, null, ids.ptr, offsets.ptr, attention.ptr, type_ids.ptr, special.ptr));
        }
    };

    /// The batch form the GPU path exists for.  `flat` = all documents back to back, `off[n_docs + 1]` byte offsets (no copy
    /// of the text is made; pinned memory is fastest, pageable memory works).  Only kept ids (u16 when the vocabulary fits)
    /// and packed offsets cross PCIe.
    pub fn encodeBatchView(self: *Self, flat: []const u8, off: []const u64, want_offsets: bool) !BatchView {
        var v: BatchView = undefined;
        const p = self.params();
        try check(cuda.tkz_encode_batch_compact(self.ctx, flat.ptr, off.ptr, off.len - 1, &p, if (want_offsets) 1 else 0, &v.r));
        return v;
    }

    /// Source-compatible batch form: one owned `Encoding` per document, exactly what a loop over Tokenizer.encode returns
    /// (src/lib.zig:109-160).  Costs six allocations per document and one per token string, as the reference does
    /// (src/encoding.zig:246-294); callers that can live with borrowed data should use `encodeBatchView`.
    pub fn encodeBatch(self: *Self, texts: []const []const u8, add_special_tokens: bool) ![]lib.Encoding {
        _ = add_special_tokens; // every post-processor of the reference is a no-op (config.zig:551-555)
        const a = self.base.allocator;
        var total: usize = 0;
        for (texts) |t| total += t.len;
        const flat = try a.alloc(u8, total);
        defer a.free(flat);
        const off = try a.alloc(u64, texts.len + 1);
        defer a.free(off);
        var pos: usize = 0;
        for (texts, 0..) |t, i| {
            off[i] = pos;
            @memcpy(flat[pos .. pos + t.len], t);
            pos += t.len;
        }
        off[texts.len] = pos;

        const view = try self.encodeBatchView(flat, off, true);

        const out = try a.alloc(lib.Encoding, texts.len);
        var built: usize = 0;
        // an allocation failure in the middle of the batch gives back everything built so far
        errdefer {
            for (out[0..built]) |*e| e.deinit();
            a.free(out);
        }
        for (0..texts.len) |d| {
            out[d] = try self.materialise(&view, d);
            built = d + 1;
        }
        return out;
    }

    /// one owned Encoding from the view (src/encoding.zig:246-294 + 385-463); frees its partial arrays on failure
    fn materialise(self: *Self, view: *const BatchView, d: usize) !lib.Encoding {
        const a = self.base.allocator;
        const n = view.slots(d);
        if (n == 0) return lib.Encoding.empty(a);
        const e_ids = try a.alloc(u32, n);
        errdefer a.free(e_ids);
        const e_type = try a.alloc(u32, n);
        errdefer a.free(e_type);
        const e_off2 = try a.alloc(u32, 2 * n);
        defer a.free(e_off2);
        const e_off = try a.alloc(lib.Offset, n);
        errdefer a.free(e_off);
        const e_spec = try a.alloc(u32, n);
        errdefer a.free(e_spec);
        const e_attn = try a.alloc(u32, n);
        errdefer a.free(e_attn);
        const e_tok = try a.alloc([]const u8, n);
        var n_tok: usize = 0;
        errdefer {
            for (e_tok[0..n_tok]) |t| a.free(t);
            a.free(e_tok);
        }
        try view.expandInto(d, e_ids, e_off2, e_attn, e_type, e_spec);
        for (0..n) |i| {
            e_off[i] = lib.Offset.init(e_off2[2 * i], e_off2[2 * i + 1]);
            // tokens[i] = the MODEL's vocab_r[id] (bpe.zig:258, wordpiece.zig:200-205); pad slots carry pad_token (encoding.zig:410,421)
            const str: []const u8 = if (e_attn[i] == 0)
                (if (self.base.padding) |pd| pd.pad_token else "[PAD]")
            else
                (self.base.model_impl.idToToken(e_ids[i]) orelse "");
            e_tok[i] = try a.dupe(u8, str);
            n_tok = i + 1;
        }
        return .{
            .allocator = a,
            .ids = e_ids,
            .type_ids = e_type,
            .tokens = e_tok,
            .offsets = e_off,
            .special_token_mask = e_spec,
            .attention_mask = e_attn,
            .words = null,
            .overflowing = &.{},
            .owns_token_strs = true,
        };
    }
};

fn componentType(root: std.json.ObjectMap, key: []const u8) ?[]const u8 {
    const v = root.get(key) orelse return null;
    if (v != .object) return null;
    const t = v.object.get("type") orelse return null;
    return if (t == .string) t.string else null;
}

fn isPunct(c: u8) bool { // config.zig:452-457
    return (c >= 33 and c <= 47) or (c >= 58 and c <= 64) or (c >= 91 and c <= 96) or (c >= 123 and c <= 126);
}
