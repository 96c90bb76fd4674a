// dff_tc_configs.h -- the instantiated configurations of the fused tcgen05 kernel: X(PN, HP, R, ATT).
// PN: padded bead-count class, HP: hidden padded to 64 / 128, R: node rows per pass, ATT: 0 = CUDA-core attention, 1 = HMMA tiles.
// Kept in one place: the build (__graft_entry__.py parses this list) compiles one translation unit per entry.
#pragma once
#define DFF_TC_CONFIGS(X) \
    X(12, 64, 60, 0) X(12, 64, 64, 0) X(32, 64, 64, 0) X(64, 64, 64, 0) \
    X(12, 128, 60, 0) X(12, 128, 64, 0) X(20, 128, 60, 0) X(32, 128, 64, 0) X(56, 128, 56, 0) \
    X(12, 64, 60, 1) X(12, 64, 64, 1) X(32, 64, 64, 1) X(64, 64, 64, 1) \
    X(12, 128, 60, 1) X(12, 128, 64, 1) X(20, 128, 60, 1) X(32, 128, 64, 1) X(56, 128, 56, 1)
