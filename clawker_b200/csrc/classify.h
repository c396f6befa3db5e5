// classify.h — host-visible launch API of classify.cu (internal to libclawker_gpu.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dev_tables.h"

namespace clk {

enum { LOADER_TMA_SW64 = 0, LOADER_TMA = 1, LOADER_DIRECT = 2 };

struct LaunchInfo { uint32_t grid, block, smem, launches; };   // launches: kernels enqueued (0 is read as 1)

// All launchers enqueue on `st` and return nullptr on success or a static error string.
const char *launch_slots(const DevTables &T, const void *d_pkts, uint32_t stride, unsigned long long n, uint32_t flags,
                         uint32_t *d_verdicts, cudaStream_t st, int loader, int stages, uint32_t *rare_q, uint32_t *rare_cnt, LaunchInfo *info);
// rest_q: device scratch of (2 + varlen_rest_cap(n)) u32 owned by this launch until it completes (leftover-frame queue)
uint32_t varlen_rest_cap(unsigned long long n);
const char *launch_varlen(const DevTables &T, const void *d_pkts, const uint32_t *d_offs16, unsigned long long n, uint32_t flags,
                          uint32_t *d_verdicts, cudaStream_t st, uint32_t *rest_q, LaunchInfo *info);
const char *launch_hdr16(const DevTables &T, const void *d_recs, const void *d_aux, unsigned long long n, uint32_t flags, uint32_t *d_verdicts,
                         cudaStream_t st, LaunchInfo *info);
const char *launch_events(const DevTables &T, const clk_event *d_ev, unsigned long long n, uint32_t flags, uint32_t *d_verdicts,
                          cudaStream_t st, LaunchInfo *info);
const char *launch_dns_extract(const DevTables &T, const void *d_pkts, const uint32_t *d_offs16, unsigned long long n,
                               const uint32_t *d_verdicts, unsigned long long *d_block_sums, unsigned long long *d_pos,
                               unsigned long long *d_total, clk_dns_update_rec *d_out, unsigned long long cap, cudaStream_t st);
const char *launch_counters_export(const DevTables &T, clk_metric *d_block, uint32_t cap_rows, cudaStream_t st);
const char *launch_counters_totals(const DevTables &T, clk_total *d_block, uint32_t rows, cudaStream_t st);
const char *launch_counters_clear(const DevTables &T, cudaStream_t st);
const char *launch_counters_import(const DevTables &T, const clk_metric *d_blocks, uint32_t n_blocks, uint32_t block_rows, cudaStream_t st);

}  // namespace clk
