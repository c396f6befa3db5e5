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
// Per caller stream: the work lists of the varlen kernel pipeline (varlen_scratch_bytes(n) bytes of device memory) and the
// side streams / events it forks onto.  Owned by the engine; a launch uses it until everything it enqueued has completed.
struct VarlenCtx {
    void *scratch = nullptr; size_t scratch_bytes = 0;
    cudaStream_t aux[3] = { nullptr, nullptr, nullptr };
    cudaEvent_t ev_head[2] = { nullptr, nullptr };
    cudaEvent_t ev_done[2][3] = { { nullptr, nullptr, nullptr }, { nullptr, nullptr, nullptr } };
};
size_t varlen_scratch_bytes(unsigned long long n);
const char *launch_varlen(const DevTables &T, const void *d_pkts, const uint32_t *d_offs16, unsigned long long n, uint32_t flags,
                          uint32_t *d_verdicts, cudaStream_t st, const VarlenCtx *ctx, LaunchInfo *info);
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
