//! build.zig -- the reference's build (build.zig:1-58) plus the static sm_100a CUDA library.
//! NOT EXECUTED in the build image (no zig toolchain).  Build the archive first:
//!     python -c 'import __graft_entry__ as g; g.build()'      -> tokenizer-zig_b200/_lib/libtokzig_b200.a
const std = @import("std");

pub fn build(b: *std.Build) void {
    const target = b.standardTargetOptions(.{});
    const optimize = b.standardOptimizeOption(.{});
    const cuda_home = b.option([]const u8, "cuda", "CUDA toolkit root") orelse "/usr/local/cuda";
    const tokzig_lib = b.option([]const u8, "tokzig-lib", "directory holding libtokzig_b200.a") orelse "../_lib";

    // module "tokenizer" exactly as the reference declares it (build.zig:8-12); src/ holds the reference sources plus
    // cuda.zig and gpu_tokenizer.zig from this directory.
    const mod = b.addModule("tokenizer", .{
        .root_source_file = b.path("src/lib.zig"),
        .target = target,
        .optimize = optimize,
        .link_libc = true,
    });
    mod.addLibraryPath(.{ .cwd_relative = tokzig_lib });
    mod.addLibraryPath(.{ .cwd_relative = b.fmt("{s}/lib64", .{cuda_home}) });
    mod.linkSystemLibrary("tokzig_b200", .{ .preferred_link_mode = .static });
    mod.linkSystemLibrary("cudart", .{}); // the CUDA runtime is linked dynamically, as the shared library of the Python / C++ harness does
    mod.linkSystemLibrary("stdc++", .{});
    mod.linkSystemLibrary("pthread", .{});

    const example = b.addExecutable(.{
        .name = "basic_tokenize",
        .root_module = b.createModule(.{
            .root_source_file = b.path("examples/basic_tokenize.zig"),
            .target = target,
            .optimize = optimize,
            .imports = &.{.{ .name = "tokenizer", .module = mod }},
        }),
    });
    b.installArtifact(example);
}
