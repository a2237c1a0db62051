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

    /// The batch form the GPU path exists for: all documents in one C-ABI call.
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

        var res: cuda.BatchResult = undefined;
        const p = self.params();
        try check(cuda.tkz_encode_batch(self.ctx, flat.ptr, off.ptr, texts.len, &p, &res));

        const out = try a.alloc(lib.Encoding, texts.len);
        errdefer a.free(out);
        for (0..texts.len) |d| {
            const lo: usize = @intCast(res.doc_tok_off.?[d]);
            const hi: usize = @intCast(res.doc_tok_off.?[d + 1]);
            const n = hi - lo;
            if (n == 0) {
                out[d] = lib.Encoding.empty(a);
                continue;
            }
            const e_ids = try a.alloc(u32, n);
            const e_type = try a.alloc(u32, n);
            const e_tok = try a.alloc([]const u8, n);
            const e_off = try a.alloc(lib.Offset, n);
            const e_spec = try a.alloc(u32, n);
            const e_attn = try a.alloc(u32, n);
            for (0..n) |i| {
                const s = lo + i;
                e_ids[i] = res.ids.?[s];
                e_type[i] = res.type_ids.?[s];
                e_off[i] = lib.Offset.init(res.offsets.?[2 * s], res.offsets.?[2 * s + 1]);
                e_spec[i] = res.special_tokens_mask.?[s];
                e_attn[i] = res.attention_mask.?[s];
                // tokens[i] == idToToken(ids[i]) always (bpe.zig:258, wordpiece.zig:200-205); pad slots carry pad_token
                const str: []const u8 = if (e_attn[i] == 0)
                    (if (self.base.padding) |pd| pd.pad_token else "[PAD]")
                else
                    (self.base.model_impl.idToToken(e_ids[i]) orelse "");
                e_tok[i] = try a.dupe(u8, str);
            }
            out[d] = .{
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
        return out;
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
