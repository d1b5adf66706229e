// Device-wide exclusive scan of an int32 array (three small launches), shared by the
// compaction and expansion kernels.  Counts are small (<= a few per element), totals fit
// int64.
#pragma once
#include "mvs_common.cuh"

// in place: a[0..n) -> exclusive prefix sums; *total (device int64) = sum.  `tile_scratch`
// must hold ceil(n/1024) int64 values.
int mvs_exclusive_scan_i32(int32_t* a, int64_t n, int64_t* tile_scratch, int64_t* total, cudaStream_t s);

// kernels one call of mvs_exclusive_scan_i32 launches (for the launch counter)
static inline int mvs_scan_launches(int64_t n) { return n == 0 ? 0 : (n <= 4096 ? 1 : ((n + 1023) / 1024 <= 2048 ? 2 : 3)); }

// grow-on-demand device buffer
int mvs_ensure(void** p, size_t* cap, size_t bytes, const char* what);
