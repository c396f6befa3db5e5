// engine.cpp — host side of libclawker_gpu.so: BPF-map-shaped tables, snapshot publication,
// batch classification entry points and counters.  Implements include/clawker_gpu.h.
//
// Threading model: every clk_map_* / clk_rules_* call takes the engine mutex and
// edits the host shadow tables only.  clk_publish() flattens them into an immutable device
// snapshot (double-buffered) — kernels receive the snapshot BY VALUE at launch, so a launch
// always sees one complete epoch, exactly like a BPF program sees old-or-new map rows, never torn.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>

#include <nccl.h>          // types and prototypes only: the library is dlopen()ed by clk_comm_init, never linked

#include "../../include/clawker_gpu.h"
#include "hdr16_pack.h"
#include "classify.h"
#include "dev_tables.h"
#include "rules.h"

using namespace clk;

namespace {

thread_local std::string tl_error;

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Snapshot {
    DevBuf ep_rows, ep_slots, ep_soa, ep6, crows, cg_slots, dns, route, cidr_pre, cidr_l1, cidr_l2_keys, cidr_l2_bits, zone, sni, names, http, http_vh;
    void release() { for (DevBuf *b : { &ep_rows, &ep_slots, &ep_soa, &ep6, &crows, &cg_slots, &dns, &route, &cidr_pre, &cidr_l1, &cidr_l2_keys, &cidr_l2_bits, &zone, &sni, &names, &http, &http_vh }) b->release(); }
};

struct MetricKeyHash {
    size_t operator()(const std::array<uint8_t, 16> &k) const { uint64_t a, b; memcpy(&a, k.data(), 8); memcpy(&b, k.data() + 8, 8); return (size_t)(mix64(a) ^ (uint64_t)mix64(b) << 32); }
};

uint32_t pow2_at_least(uint64_t n) { uint32_t c = 16; while (c < n) c <<= 1; return c; }

}  // namespace

struct clk_engine {
    std::mutex mu;
    std::string err;
    bool host_only = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = { nullptr, nullptr, nullptr };
    cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool ev_recorded = false;   // ev_recorded: both events have been recorded at least once
    uint32_t max_containers = 256, max_dns = 16384, max_routes = 8192, max_metrics = 16384, max_cidrs = 1u << 20, reserved_sms = 0;

    // host shadow tables (the BPF maps)
    std::unordered_map<uint64_t, clk_container_config> container;
    std::unordered_map<uint64_t, uint8_t> bypass;
    std::unordered_map<uint32_t, clk_dns_entry> dns;
    std::unordered_map<uint64_t, uint16_t> route;                       // (hash << 16 | port) -> envoy_port
    std::unordered_map<uint32_t, clk_endpoint> endpoint;
    std::map<std::array<uint8_t, 16>, clk_endpoint> endpoint6;
    std::map<std::pair<uint32_t, uint32_t>, uint8_t> cidr;              // (addr, mask)
    CompiledRules compiled;                                              // zones / snis / mappings
    bool dirty = true;

    // stable cgroup -> ContRow index (metric keys reference it)
    std::unordered_map<uint64_t, uint32_t> crow_of;
    std::vector<uint64_t> crow_cg;

    // device
    Snapshot snap[2]; int active = 0; bool published = false;
    DevTables T;                                                         // active snapshot
    DevBuf m_keys, m_counts, m_dropped;      // this GPU's metrics_map (kernels write)
    DevBuf g_keys, g_counts, g_dropped;      // merged view after the all-gather (clk_counters_import)
    uint32_t g_mask = 0;                       // its slot mask: sized for n_blocks full tables at load <= 1/2
    DevBuf work_rare[3];                       // rare / leftover-frame queues of the host pipeline lanes
    std::map<cudaStream_t, DevBuf> scratch;    // the same for the *_dev entries: one per caller stream, so launches in flight on
                                               // different streams of one engine never share a queue
    struct VarlenState { DevBuf buf; VarlenCtx ctx; };
    std::map<cudaStream_t, VarlenState> varlen; // work lists + side streams / events of the varlen kernel pipeline, per caller stream
    DevBuf work_pkts[3], work_offs[3], work_verd[3], scan_sums, scan_pos, scan_total, upd_out, ctr_block;
    // multi-GPU (clk_comm_*): one NCCL communicator per engine, engine-owned gather buffers
    ncclComm_t comm = nullptr; int comm_rank = 0, comm_world = 1;
    DevBuf tot_local, tot_all, met_local, met_all;
    clk_stats stats;
    int loader = LOADER_TMA_SW64;
    int stages = 1;
    bool stage_tables = true;
    uint64_t chunk_pkts = 1u << 21;

    clk_engine() { memset(&T, 0, sizeof T); memset(&stats, 0, sizeof stats); }
};

namespace {

int fail(clk_engine *e, int code, const char *fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    tl_error = buf;
    if (e) e->err = buf;
    return code;
}
#define CUDA_TRY(e, call) do { cudaError_t _c = (call); if (_c != cudaSuccess) return fail((e), CLK_EIO, "%s: %s", #call, cudaGetErrorString(_c)); } while (0)
// Every GPU entry point runs on the engine's device whatever the calling thread's current device is (a host with
// one engine per GPU calls from arbitrary threads; streams and buffers of another device are "invalid resource handle").
#define NEED_GPU(e) do { if ((e)->host_only) return fail((e), CLK_ENOSYS, "engine was created host-only (no CUDA device): classification requires the GPU"); \
                         cudaError_t _d = cudaSetDevice((e)->device); if (_d != cudaSuccess) return fail((e), CLK_EIO, "cudaSetDevice(%d): %s", (e)->device, cudaGetErrorString(_d)); \
                         (void)cudaGetLastError(); /* a stale error left by another library on this thread is not ours to report */ } while (0)


// ---- NCCL, loaded on demand --------------------------------------------------------------------------------------
// libclawker_gpu.so does not link NCCL: a single-GPU host never needs it, and inside a process that already carries
// an NCCL (PyTorch bundles its own libnccl.so.2) dlopen() by SONAME hands back that very copy, so there are never two.
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi &nccl()
{
    static NcclApi api; static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = { getenv("CLK_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) { if (n && *n && (api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (!api.h) { api.err = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) { api.err = "libnccl.so.2 lacks an expected symbol"; api.h = nullptr; }
    });
    return api;
}
void comm_release(clk_engine *e)
{
    if (e->comm && nccl().h) nccl().CommDestroy(e->comm);
    e->comm = nullptr; e->comm_world = 1; e->comm_rank = 0;
    for (DevBuf *b : { &e->tot_local, &e->tot_all, &e->met_local, &e->met_all }) b->release();
}
#define NCCL_TRY(e, call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) return fail((e), CLK_EIO, "%s: %s", #call, nccl().GetErrorString(_r)); } while (0)

struct MapSpec { size_t klen, vlen; };
const MapSpec kSpecs[CLK_MAP__COUNT] = {
    { 8, sizeof(clk_container_config) }, { 8, 1 }, { 4, sizeof(clk_dns_entry) }, { sizeof(clk_route_key), sizeof(clk_route_val) },
    { sizeof(clk_metric_key), 8 }, { 4, sizeof(clk_endpoint) }, { 16, sizeof(clk_endpoint) }, { sizeof(clk_cidr), 1 },
};

uint64_t route_k(const clk_route_key *k) { return ((uint64_t)k->domain_hash << 16) | k->dst_port; }

bool is_cidr_mask(uint32_t mask_nbo)
{
    uint32_t h = __builtin_bswap32(mask_nbo);
    return ((~h) & ((~h) + 1)) == 0;          // ~h is 2^k - 1
}
int mask_plen(uint32_t mask_nbo) { return __builtin_popcount(mask_nbo); }

std::string to_lower_ascii(std::string s) { for (auto &c : s) if (c >= 'A' && c <= 'Z') c = (char)(c + 32); return s; }

uint64_t name_key_of(const std::string &s)
{
    uint32_t h1 = 0, h2 = 0, p1 = 1, p2 = 1;
    for (unsigned char c : s) { if (c >= 'A' && c <= 'Z') c = (unsigned char)(c + 32); h1 += c * p1; h2 += c * p2; p1 *= NAME_B1; p2 *= NAME_B2; }
    return name_key(h1, h2, (uint32_t)s.size());
}

// Build one name table (slots + blob entries appended to `blob`).
void build_name_table(const std::vector<std::string> &names, const std::vector<uint32_t> &vals, std::vector<NameSlot> &slots,
                      std::vector<uint8_t> &blob)
{
    uint32_t cap = pow2_at_least((uint64_t)names.size() * 4 + 2);          // load <= 1/4, even home slots (dev_tables.h name_home)
    slots.assign(cap, NameSlot{ 0, 0, 0 });
    for (size_t i = 0; i < names.size(); i++) {
        const std::string &s = names[i];
        if (s.size() > 255) continue;
        while (blob.size() % 16) blob.push_back(0);                        // entries are 16-byte aligned and zero padded (name_load64)
        const uint32_t off16 = (uint32_t)(blob.size() / 16);
        for (unsigned char c : s) blob.push_back((c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c);
        while (blob.size() % 16) blob.push_back(0);
        uint64_t key = name_key_of(s);
        uint32_t h = name_home(key, cap - 1);
        while (slots[h].key) h = (h + 1) & (cap - 1);
        slots[h] = NameSlot{ key, vals[i], ((uint32_t)s.size() << 24) | off16 };
    }
}

template <typename V> cudaError_t upload(DevBuf &b, const std::vector<V> &v, cudaStream_t st)
{
    if (v.empty()) return cudaSuccess;
    cudaError_t e = b.ensure(v.size() * sizeof(V));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice, st);
}

uint32_t crow_index(clk_engine *e, uint64_t cg)
{
    auto it = e->crow_of.find(cg);
    if (it != e->crow_of.end()) return it->second;
    if (e->crow_cg.size() >= CLK_MAX_CROWS) return 0xffffffffu;
    uint32_t idx = (uint32_t)e->crow_cg.size();
    e->crow_cg.push_back(cg); e->crow_of[cg] = idx;
    return idx;
}

// The cgroup -> row index only grows while cgroups come and go (every container start brings a new cgroup id), and a
// metric key carries the row index, so rows cannot simply be renumbered.  When dead rows pile up — cgroups that are in
// neither container_map nor bypass_map — this reads the metric rows back (they name cgroups by id), keeps the cgroups
// that are live or still own a metric row, renumbers, and hands the rows back for re-import under the new numbering
// once the new snapshot is active.  `pending` receives the rows to re-import (empty when nothing was compacted).
int compact_crows_locked(clk_engine *e, std::vector<clk_metric> &pending, bool &compacted)
{
    compacted = false;
    size_t live_now = 0;
    for (uint64_t cg : e->crow_cg) if (e->container.count(cg) || e->bypass.count(cg)) live_now++;
    size_t incoming = 0;
    for (auto &kv : e->container) if (!e->crow_of.count(kv.first)) incoming++;
    for (auto &kv : e->bypass) if (!e->crow_of.count(kv.first) && !e->container.count(kv.first)) incoming++;
    const size_t dead = e->crow_cg.size() - live_now;
    if (dead < 64 && e->crow_cg.size() + incoming <= 2ull * e->max_containers && e->crow_cg.size() + incoming <= CLK_MAX_CROWS) return CLK_OK;
    std::set<uint64_t> keep;
    if (e->published && e->T.m_keys) {                  // which cgroups still own metric rows (old snapshot: old numbering)
        const uint32_t rows = e->T.m_mask + 2;
        CUDA_TRY(e, e->ctr_block.ensure((size_t)rows * sizeof(clk_metric)));
        const char *err = launch_counters_export(e->T, (clk_metric *)e->ctr_block.p, rows, e->stream);
        if (err) return fail(e, CLK_EIO, "counters export: %s", err);
        std::vector<clk_metric> host(rows);
        CUDA_TRY(e, cudaMemcpyAsync(host.data(), e->ctr_block.p, (size_t)rows * sizeof(clk_metric), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(e, cudaStreamSynchronize(e->stream));
        const uint64_t valid = std::min<uint64_t>(host[0].count, rows - 1);
        for (uint64_t i = 1; i <= valid; i++) if (host[i].count) { pending.push_back(host[i]); keep.insert(host[i].key.cgroup_id); }
    }
    std::vector<uint64_t> order;
    for (uint64_t cg : e->crow_cg) if (e->container.count(cg) || e->bypass.count(cg) || keep.count(cg)) order.push_back(cg);
    if (order.size() == e->crow_cg.size()) { pending.clear(); return CLK_OK; }      // nothing to reclaim
    e->crow_cg = order; e->crow_of.clear();
    for (size_t i = 0; i < order.size(); i++) e->crow_of[order[i]] = (uint32_t)i;
    compacted = true;
    return CLK_OK;
}

int publish_locked(clk_engine *e)
{
    NEED_GPU(e);
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());              // no kernel may still read the buffers we are about to overwrite
    Snapshot &S = e->snap[1 - e->active];
    DevTables T; memset(&T, 0, sizeof T);
    cudaStream_t st = e->stream;

    // ---- cgroup rows (stable indices; dead rows are reclaimed, see compact_crows_locked) ----------
    std::vector<clk_metric> reimport; bool compacted = false;
    { int rc = compact_crows_locked(e, reimport, compacted); if (rc) return rc; }
    for (auto &kv : e->container) if (crow_index(e, kv.first) == 0xffffffffu) return fail(e, CLK_E2BIG, "more than %u distinct cgroups with live metric rows", CLK_MAX_CROWS);
    for (auto &kv : e->bypass)    if (crow_index(e, kv.first) == 0xffffffffu) return fail(e, CLK_E2BIG, "more than %u distinct cgroups with live metric rows", CLK_MAX_CROWS);
    std::vector<ContRow> crows(e->crow_cg.size());
    for (size_t i = 0; i < crows.size(); i++) {
        ContRow &r = crows[i]; memset(&r, 0, sizeof r);
        r.cgroup_id = e->crow_cg[i];
        auto c = e->container.find(r.cgroup_id);
        if (c != e->container.end()) { r.flags |= CLK_ROW_ENFORCED; r.cfg = c->second; }
        auto b = e->bypass.find(r.cgroup_id);
        if (b != e->bypass.end() && b->second == 1) r.flags |= CLK_ROW_BYPASS;
    }
    std::vector<uint32_t> cg_slots(pow2_at_least(crows.size() * 2 + 1), 0);
    for (size_t i = 0; i < crows.size(); i++) {
        uint32_t h = mix64(crows[i].cgroup_id) & (uint32_t)(cg_slots.size() - 1);
        while (cg_slots[h]) h = (h + 1) & (uint32_t)(cg_slots.size() - 1);
        cg_slots[h] = (uint32_t)i + 1;
    }
    // ---- endpoint rows ----------------------------------------------------------------------------
    std::vector<EpRow> rows; std::vector<Ep6Key> ep6_list, ep6;
    auto fill = [&](EpRow &r, const clk_endpoint &ep) {
        r.uid = ep.uid; r.cgroup_id = ep.cgroup_id; r.so_mark = ep.so_mark; r.crow = 0xffffffffu;
        auto it = e->crow_of.find(ep.cgroup_id);
        if (it != e->crow_of.end()) { r.crow = it->second; r.flags = crows[it->second].flags; r.cfg = crows[it->second].cfg; }
    };
    for (auto &kv : e->endpoint) { EpRow r; memset(&r, 0, sizeof r); r.ip = kv.first; fill(r, kv.second); rows.push_back(r); }
    const size_t n_ep4 = rows.size();
    std::vector<uint2> ep_slots(pow2_at_least(n_ep4 * 4 + 1), uint2{ 0, 0 });      // load <= 1/4: probe loops stay ~1 iteration
    for (size_t i = 0; i < n_ep4; i++) {
        uint32_t h = hash_a(rows[i].ip) >> shift_for((uint32_t)ep_slots.size());
        while (ep_slots[h].y) h = (h + 1) & (uint32_t)(ep_slots.size() - 1);
        ep_slots[h] = uint2{ rows[i].ip, (uint32_t)i + 1 };
    }
    for (auto &kv : e->endpoint6) {
        EpRow r; memset(&r, 0, sizeof r); fill(r, kv.second);
        Ep6Key k; memset(&k, 0, sizeof k); memcpy(k.a, kv.first.data(), 16); k.row = (uint32_t)rows.size();
        rows.push_back(r); ep6_list.push_back(k);
    }
    // ---- dns_cache / route_map --------------------------------------------------------------------
    if (!ep6_list.empty()) {                                   // IPv6 endpoints: open addressing at load <= 1/4
        Ep6Key empty; memset(&empty, 0, sizeof empty); empty.row = 0xffffffffu;
        ep6.assign(pow2_at_least(ep6_list.size() * 4 + 1), empty);
        const uint32_t mask = (uint32_t)ep6.size() - 1, sh = shift_for((uint32_t)ep6.size());
        for (auto &k : ep6_list) {
            uint32_t hh = hash_a(k.a[0] ^ k.a[1] ^ k.a[2] ^ k.a[3]) >> sh;
            while (ep6[hh].row != 0xffffffffu) hh = (hh + 1) & mask;
            ep6[hh] = k;
        }
    }
    // 2-choice bucketed tables (dev_tables.h): grow until every key fits one of its two buckets
    std::vector<uint4> dns_b, route_b;
    auto slot_empty = [](const uint4 &b, int k) { return k == 0 ? (b.x & b.y) == 0xffffffffu : (b.z & b.w) == 0xffffffffu; };
    auto slot_put = [](uint4 &b, int k, uint64_t s) { if (k == 0) { b.x = (uint32_t)s; b.y = (uint32_t)(s >> 32); } else { b.z = (uint32_t)s; b.w = (uint32_t)(s >> 32); } };
    // items: (slot word, (hash_a, hash_b) of the key).  Invariant the lookup relies on: a key sits in its
    // second bucket only while its first bucket is full.  One level of displacement (move a resident whose
    // FIRST bucket this is over to its second bucket) keeps tables at ~0.4 load.
    auto build_buckets = [&](std::vector<uint4> &out, const std::vector<std::pair<uint64_t, std::pair<uint32_t, uint32_t>>> &items) {
        uint32_t nb = pow2_at_least((uint64_t)items.size() * 5 / 6 + 1);
        for (;; nb *= 2) {
            const uint32_t sh = shift_for(nb);
            out.assign(nb, uint4{ 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu });
            std::vector<int> owner((size_t)nb * 2, -1);                  // item index per slot
            auto place = [&](uint32_t b, int idx) -> bool {
                for (int k = 0; k < 2; k++) if (slot_empty(out[b], k)) { slot_put(out[b], k, items[idx].first); owner[(size_t)b * 2 + k] = idx; return true; }
                return false;
            };
            bool ok = true;
            for (int idx = 0; idx < (int)items.size() && ok; idx++) {
                const uint32_t b1 = items[idx].second.first >> sh, b2 = items[idx].second.second >> sh;
                if (place(b1, idx) || place(b2, idx)) continue;
                bool moved = false;
                for (uint32_t b : { b1, b2 }) {
                    for (int k = 0; k < 2 && !moved; k++) {
                        int j = owner[(size_t)b * 2 + k];
                        if (j < 0 || (items[j].second.first >> sh) != b) continue;          // only residents in their first bucket may move
                        uint32_t jb2 = items[j].second.second >> sh;
                        if (jb2 != b && place(jb2, j)) { slot_put(out[b], k, items[idx].first); owner[(size_t)b * 2 + k] = idx; moved = true; }
                    }
                    if (moved) break;
                }
                ok = moved;
            }
            if (ok) break;
        }
    };
    if (!e->dns.empty()) {
        std::vector<std::pair<uint64_t, std::pair<uint32_t, uint32_t>>> items;
        for (auto &kv : e->dns) {
            uint64_t s = (uint64_t)kv.first | ((uint64_t)kv.second.domain_hash << 32);
            if (kv.first == 0xffffffffu) { T.dns_ffff_present = 1; T.dns_ffff_hash = kv.second.domain_hash; continue; }
            items.push_back({ s, { hash_a(kv.first), hash_b(kv.first) } });
        }
        if (!items.empty()) build_buckets(dns_b, items);
    }
    if (!e->route.empty()) {
        std::vector<std::pair<uint64_t, std::pair<uint32_t, uint32_t>>> items;
        for (auto &kv : e->route) {
            uint32_t hash = (uint32_t)(kv.first >> 16), port = (uint32_t)(kv.first & 0xffff), f = route_fold(hash, port);
            items.push_back({ (kv.first << 16) | kv.second, { hash_a(f), hash_b(f) } });
        }
        build_buckets(route_b, items);
    }
    // ---- CIDR directory (A12) ---------------------------------------------------------------------
    std::vector<uint32_t> l1, l2_keys, l2_bits, pre;
    uint32_t n_cidr = 0;
    if (!e->cidr.empty()) {
        l1.assign((1u << 24) / 16, 0); pre.assign(65536 / 32, 0);
        std::unordered_map<uint32_t, std::array<uint32_t, 8>> leaves;
        for (auto &kv : e->cidr) {
            uint32_t addr = kv.first.first, mask = kv.first.second;
            if ((addr & mask) != addr) continue;           // is_in_subnet can never be true for it (common.h:220-223)
            n_cidr++;
            int plen = mask_plen(mask);
            uint32_t base = __builtin_bswap32(addr);       // host-order network
            {   // /16 prefilter
                uint32_t p0 = base >> 16, pc = plen >= 16 ? 1u : 1u << (16 - plen);
                for (uint32_t p = p0; p < p0 + pc; p++) pre[p >> 5] |= 1u << (p & 31);
            }
            if (plen <= 24) {
                uint32_t b0 = base >> 8, cnt = 1u << (24 - plen);
                for (uint32_t b = b0; b < b0 + cnt; b++) l1[b >> 4] |= 1u << ((b & 15) * 2);
            } else {
                uint32_t blk = base >> 8, h0 = base & 255u, cnt = 1u << (32 - plen);
                l1[blk >> 4] |= 2u << ((blk & 15) * 2);
                auto &leaf = leaves[blk];
                for (uint32_t hst = h0; hst < h0 + cnt; hst++) leaf[hst >> 5] |= 1u << (hst & 31);
            }
        }
        uint32_t cap = pow2_at_least(leaves.size() * 2 + 1);
        l2_keys.assign(cap, 0); l2_bits.assign((size_t)cap * 8, 0);
        for (auto &kv : leaves) {
            uint32_t h = hash_a(kv.first) >> shift_for(cap);
            while (l2_keys[h]) h = (h + 1) & (cap - 1);
            l2_keys[h] = kv.first + 1;
            memcpy(&l2_bits[(size_t)h * 8], kv.second.data(), 32);
        }
        if (n_cidr == 0) { l1.clear(); l2_keys.clear(); l2_bits.clear(); pre.clear(); }
    }
    // ---- name tables --------------------------------------------------------------------------------
    std::vector<NameSlot> zone_slots, sni_slots; std::vector<uint8_t> blob;
    if (!e->compiled.zones.empty()) {
        std::vector<uint32_t> vals; for (auto &z : e->compiled.zones) vals.push_back(domain_hash(z.data(), z.size()));
        build_name_table(e->compiled.zones, vals, zone_slots, blob);
    }
    if (!e->compiled.snis.empty()) {
        std::vector<uint32_t> vals(e->compiled.snis.size()); for (size_t i = 0; i < vals.size(); i++) vals[i] = (uint32_t)i;
        build_name_table(e->compiled.snis, vals, sni_slots, blob);
    }
    // F3: virtual hosts of the plaintext-HTTP chain.  Domain forms (envoy.go:932-944): "d" exact, "*.d" suffix
    // wildcard (stored ".d"), "d:*" prefix wildcard (stored "d:"); the first virtual host that lists a form owns it.
    std::vector<NameSlot> http_slots; std::vector<uint32_t> http_vh;
    if (!e->compiled.http.empty()) {
        std::vector<std::string> names; std::vector<uint32_t> vals; std::set<std::pair<uint32_t, std::string>> seen;
        for (size_t v = 0; v < e->compiled.http.size(); v++) {
            for (const std::string &d : e->compiled.http[v].domains) {
                if (d == "*" || d.empty()) continue;                    // the catch-all is the implicit deny_all host
                uint32_t kind; std::string str;
                if (d[0] == '*') { kind = HTTP_K_SUFFIX; str = d.substr(1); }
                else if (d.back() == '*') { kind = HTTP_K_PREFIX; str = d.substr(0, d.size() - 1); }
                else { kind = HTTP_K_EXACT; str = d; }
                if (str.size() > 255 || (kind == HTTP_K_SUFFIX && str[0] != '.') || (kind == HTTP_K_PREFIX && str.back() != ':'))
                    return fail(e, CLK_EINVAL, "virtual-host domain \"%s\": unsupported form", d.c_str());
                if (!seen.insert({ kind, to_lower_ascii(str) }).second) continue;
                names.push_back(str); vals.push_back((uint32_t)v | (kind << 28));
            }
            const auto &routes = e->compiled.http[v].routes;
            if (routes.size() > 255) return fail(e, CLK_E2BIG, "virtual host %zu: %zu routes > 255", v, routes.size());
            http_vh.push_back((uint32_t)blob.size());
            blob.push_back((uint8_t)routes.size());
            for (const auto &r : routes) {
                if (r.prefix.size() > 255) return fail(e, CLK_E2BIG, "path prefix longer than 255 bytes");
                blob.push_back(r.allow ? 1 : 0); blob.push_back((uint8_t)r.prefix.size());
                blob.insert(blob.end(), r.prefix.begin(), r.prefix.end());
            }
        }
        build_name_table(names, vals, http_slots, blob);
    }
    if (blob.size() >= (1ull << 28)) return fail(e, CLK_E2BIG, "name tables need %zu bytes of strings (limit 256 MiB)", blob.size());
    // ---- upload -------------------------------------------------------------------------------------
    CUDA_TRY(e, upload(S.crows, crows, st));       CUDA_TRY(e, upload(S.cg_slots, cg_slots, st));
    // the fields the fast path needs, packed as three 16-byte words per row (48 B stride: conflict-free LDS.128 for 8 rows)
    std::vector<uint4> rowq(std::max<size_t>(rows.size(), 1) * 3, uint4{ 0, 0, 0, 0 });
    for (size_t i = 0; i < rows.size(); i++) {
        const EpRow &r = rows[i];
        rowq[3 * i + 0] = uint4{ r.uid, r.so_mark, r.flags, r.crow };
        rowq[3 * i + 1] = uint4{ r.cfg.envoy_ip, r.cfg.coredns_ip, r.cfg.gateway_ip, r.cfg.host_proxy_ip };
        rowq[3 * i + 2] = uint4{ r.cfg.net_addr, r.cfg.net_mask, (uint32_t)r.cfg.host_proxy_port | ((uint32_t)r.cfg.egress_port << 16), 0 };
    }
    CUDA_TRY(e, upload(S.ep_rows, rows, st));      CUDA_TRY(e, upload(S.ep_slots, ep_slots, st));
    CUDA_TRY(e, upload(S.ep_soa, rowq, st));       CUDA_TRY(e, upload(S.ep6, ep6, st));
    CUDA_TRY(e, upload(S.dns, dns_b, st));         CUDA_TRY(e, upload(S.route, route_b, st));
    CUDA_TRY(e, upload(S.cidr_pre, pre, st));
    CUDA_TRY(e, upload(S.cidr_l1, l1, st));        CUDA_TRY(e, upload(S.cidr_l2_keys, l2_keys, st));
    CUDA_TRY(e, upload(S.cidr_l2_bits, l2_bits, st));
    CUDA_TRY(e, upload(S.zone, zone_slots, st));   CUDA_TRY(e, upload(S.sni, sni_slots, st));
    CUDA_TRY(e, upload(S.names, blob, st));
    CUDA_TRY(e, upload(S.http, http_slots, st));   CUDA_TRY(e, upload(S.http_vh, http_vh, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));

    T.ep_rows = rows.empty() ? nullptr : (const EpRow *)S.ep_rows.p; T.n_ep_rows = (uint32_t)rows.size();
    if (n_ep4) { T.ep_slots = (const uint2 *)S.ep_slots.p; T.ep_mask = (uint32_t)ep_slots.size() - 1; T.ep_shift = shift_for((uint32_t)ep_slots.size()); }
    T.ep_rowq = (const uint4 *)S.ep_soa.p;
    T.ep6 = ep6.empty() ? nullptr : (const Ep6Key *)S.ep6.p; T.n_ep6 = (uint32_t)ep6_list.size();
    if (!ep6.empty()) { T.ep6_mask = (uint32_t)ep6.size() - 1; T.ep6_shift = shift_for((uint32_t)ep6.size()); }
    T.crows = crows.empty() ? nullptr : (const ContRow *)S.crows.p; T.n_crows = (uint32_t)crows.size();
    T.cg_slots = (const uint32_t *)S.cg_slots.p; T.cg_mask = (uint32_t)cg_slots.size() - 1;
    if (!dns_b.empty()) { T.dns_b = (const uint4 *)S.dns.p; T.dns_shift = shift_for((uint32_t)dns_b.size()); }
    if (!route_b.empty()) { T.route_b = (const uint4 *)S.route.p; T.route_shift = shift_for((uint32_t)route_b.size()); }
    T.n_cidr = n_cidr;
    if (n_cidr) { T.cidr_pre = (const uint32_t *)S.cidr_pre.p; T.cidr_l1 = (const uint32_t *)S.cidr_l1.p; T.cidr_l2_keys = (const uint32_t *)S.cidr_l2_keys.p; T.cidr_l2_bits = (const uint32_t *)S.cidr_l2_bits.p; T.cidr_l2_mask = (uint32_t)l2_keys.size() - 1; T.cidr_l2_shift = shift_for((uint32_t)l2_keys.size()); }
    if (!zone_slots.empty()) { T.zone_slots = (const NameSlot *)S.zone.p; T.zone_mask = (uint32_t)zone_slots.size() - 1; }
    if (!sni_slots.empty()) { T.sni_slots = (const NameSlot *)S.sni.p; T.sni_mask = (uint32_t)sni_slots.size() - 1; }
    for (auto &z : e->compiled.zones) T.zone_max_labels = std::max<uint32_t>(T.zone_max_labels, 1u + (uint32_t)std::count(z.begin(), z.end(), '.'));
    for (auto &z : e->compiled.snis) if (!z.empty() && z[0] == '.') T.sni_max_dots = std::max<uint32_t>(T.sni_max_dots, (uint32_t)std::count(z.begin(), z.end(), '.'));
    T.names = (const uint8_t *)S.names.p;
    if (!http_vh.empty()) {
        T.n_http = (uint32_t)http_vh.size(); T.http_vh = (const uint32_t *)S.http_vh.p;
        if (!http_slots.empty()) { T.http_slots = (const NameSlot *)S.http.p; T.http_mask = (uint32_t)http_slots.size() - 1; }
    }
    // distinct CoreDNS addresses (DNS-answer stage)
    for (auto &kv : e->container) {
        bool seen = false;
        for (uint32_t i = 0; i < T.n_coredns; i++) seen |= T.coredns_ips[i] == kv.second.coredns_ip;
        if (seen) continue;
        if (T.n_coredns >= 8) return fail(e, CLK_E2BIG, "more than 8 distinct coredns_ip values in container_map (DNS answers to a 9th CoreDNS would go unseen)");
        T.coredns_ips[T.n_coredns++] = kv.second.coredns_ip;
    }
    std::sort(T.coredns_ips, T.coredns_ips + T.n_coredns);
    T.m_keys = (unsigned long long *)e->m_keys.p; T.m_counts = (unsigned long long *)e->m_counts.p;
    T.m_mask = e->T.m_mask; T.m_dropped = (unsigned long long *)e->m_dropped.p; T.m_max = e->max_metrics; T.reserved_sms = e->reserved_sms;
    // shared-memory staging budget: endpoints <= 256 rows, dns <= 32 KiB, route <= 8 KiB
    if (e->stage_tables) {
        if (n_ep4 && rows.size() <= 256) { T.stage_ep_bytes = (uint32_t)ep_slots.size() * 8; T.stage_soa_bytes = (uint32_t)rowq.size() * 16; }
        if (!dns_b.empty() && dns_b.size() * 16 <= 32768) T.stage_dns_bytes = (uint32_t)dns_b.size() * 16;
        if (!route_b.empty() && route_b.size() * 16 <= 8192) T.stage_route_bytes = (uint32_t)route_b.size() * 16;
        if (n_cidr) T.stage_pre_bytes = 8192;
        if (!ep6.empty() && ep6.size() * 32 <= 8192) T.stage_ep6_bytes = (uint32_t)ep6.size() * 32;
    }
    e->T = T; e->active = 1 - e->active; e->published = true; e->dirty = false;
    if (compacted) {                                    // the metric rows again, keyed by the new row numbers
        unsigned long long ctr[2] = { 0, 0 };            // [0] increments dropped so far survives the rebuild
        CUDA_TRY(e, cudaMemcpy(ctr, e->m_dropped.p, 16, cudaMemcpyDeviceToHost));
        const char *err = launch_counters_clear(e->T, st);
        if (!err && !reimport.empty()) {
            std::vector<clk_metric> block(reimport.size() + 1);
            memset(&block[0], 0, sizeof(clk_metric)); block[0].count = reimport.size();
            std::copy(reimport.begin(), reimport.end(), block.begin() + 1);
            CUDA_TRY(e, e->ctr_block.ensure(block.size() * sizeof(clk_metric)));
            CUDA_TRY(e, cudaMemcpyAsync(e->ctr_block.p, block.data(), block.size() * sizeof(clk_metric), cudaMemcpyHostToDevice, st));
            DevTables U = e->T; U.m_max = 0xffffffffu;   // these rows were admitted before: the max_entries gate does not apply again
            err = launch_counters_import(U, (const clk_metric *)e->ctr_block.p, 1, (uint32_t)block.size(), st);
        }
        if (err) return fail(e, CLK_EIO, "re-import of the metric rows after cgroup-row compaction: %s", err);
        CUDA_TRY(e, cudaStreamSynchronize(st));
        ctr[1] = reimport.size();                        // the key counter the max_entries gate reads
        CUDA_TRY(e, cudaMemcpy(e->m_dropped.p, ctr, 16, cudaMemcpyHostToDevice));
    }
    return CLK_OK;
}

// Work lists, side streams and events of the varlen pipeline for launches enqueued on `st` (created on first use).
int varlen_ctx(clk_engine *e, cudaStream_t st, uint64_t n, const VarlenCtx **out)
{
    clk_engine::VarlenState &v = e->varlen[st];
    if (!v.ctx.aux[0]) {
        for (int a = 0; a < 3; a++) CUDA_TRY(e, cudaStreamCreateWithFlags(&v.ctx.aux[a], cudaStreamNonBlocking));
        for (int p = 0; p < 2; p++) {
            CUDA_TRY(e, cudaEventCreateWithFlags(&v.ctx.ev_head[p], cudaEventDisableTiming));
            for (int a = 0; a < 3; a++) CUDA_TRY(e, cudaEventCreateWithFlags(&v.ctx.ev_done[p][a], cudaEventDisableTiming));
        }
    }
    CUDA_TRY(e, v.buf.ensure(varlen_scratch_bytes(n)));
    v.ctx.scratch = v.buf.p; v.ctx.scratch_bytes = v.buf.cap;
    *out = &v.ctx;
    return CLK_OK;
}

int ensure_published(clk_engine *e)
{
    if (!e->published || e->dirty) return publish_locked(e);
    return CLK_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
// Exception barrier of the C ABI (include/clawker_gpu.h: "no call aborts or throws"): every export is a function-try-block,
// so a std::bad_alloc / length_error from the host containers becomes CLK_ENOMEM / CLK_EIO instead of unwinding into cgo.
#define CLK_GUARD_INT(e) catch (const std::bad_alloc &) { return fail((e), CLK_ENOMEM, "out of memory"); } \
    catch (const std::exception &x) { return fail((e), CLK_EIO, "internal error: %s", x.what()); } catch (...) { return fail((e), CLK_EIO, "internal error"); }
#define CLK_GUARD_VAL(v) catch (...) { return (v); }
#define CLK_GUARD_VOID catch (...) { }

extern "C" {

uint32_t clk_abi_version(void) { return CLK_ABI_VERSION; }

const char *clk_strerror(int code)
{
    switch (code) {
    case CLK_OK: return "ok"; case CLK_ENOENT: return "key does not exist"; case CLK_EIO: return "CUDA error";
    case CLK_E2BIG: return "map full"; case CLK_ENOMEM: return "out of memory"; case CLK_EEXIST: return "exists";
    case CLK_EINVAL: return "invalid argument"; case CLK_ENOSYS: return "not available (host-only engine)";
    default: return "unknown error";
    }
}

const char *clk_last_error(const clk_engine *e) { return e ? e->err.c_str() : tl_error.c_str(); }

int clk_engine_create(const clk_cfg *cfg, clk_engine **out)
try {
    if (!out) return fail(nullptr, CLK_EINVAL, "out is NULL");
    *out = nullptr;
    if (cfg && cfg->abi_version != CLK_ABI_VERSION) return fail(nullptr, CLK_EINVAL, "abi_version %u != %u", cfg->abi_version, CLK_ABI_VERSION);
    clk_engine *e = new (std::nothrow) clk_engine();
    if (!e) return fail(nullptr, CLK_ENOMEM, "out of memory");
    if (cfg) {
        if (cfg->max_containers) e->max_containers = cfg->max_containers;
        if (cfg->max_dns_cache) e->max_dns = cfg->max_dns_cache;
        if (cfg->max_routes) e->max_routes = cfg->max_routes;
        if (cfg->max_metrics) e->max_metrics = cfg->max_metrics;
        if (cfg->max_cidrs) e->max_cidrs = cfg->max_cidrs;
        e->host_only = (cfg->flags & 1u) != 0;
        e->reserved_sms = cfg->reserved_sms;
    }
    if (const char *s = getenv("CLK_LOADER")) e->loader = atoi(s);
    if (const char *s = getenv("CLK_STAGE")) e->stage_tables = atoi(s) != 0;
    if (const char *s = getenv("CLK_STAGES")) e->stages = atoi(s);
    if (const char *s = getenv("CLK_MINB")) e->stages |= atoi(s) << 4;
    if (const char *s = getenv("CLK_CHUNK_PKTS")) { long v = atol(s); if (v >= 1024) e->chunk_pkts = (uint64_t)v; }
    if (!e->host_only) {
        int dev = cfg ? cfg->device : -1;
        cudaError_t c;
        if (dev < 0) { c = cudaGetDevice(&dev); if (c != cudaSuccess) { int r = fail(nullptr, CLK_EIO, "no CUDA device: %s", cudaGetErrorString(c)); delete e; return r; } }
        e->device = dev;
        int n_sm = 0;
        if ((c = cudaSetDevice(dev)) != cudaSuccess || (c = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (c = cudaEventCreate(&e->ev0)) != cudaSuccess || (c = cudaEventCreate(&e->ev1)) != cudaSuccess ||
            (c = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) {
            int r = fail(nullptr, CLK_EIO, "CUDA init on device %d failed: %s", dev, cudaGetErrorString(c)); delete e; return r;
        }
        for (int i = 0; i < 3; i++) cudaStreamCreateWithFlags(&e->pipe[i], cudaStreamNonBlocking);
        e->stats.sm_count = (uint32_t)n_sm;
        uint32_t slots = pow2_at_least((uint64_t)e->max_metrics * 2);
        if (e->m_keys.ensure((size_t)slots * 8) != cudaSuccess || e->m_counts.ensure((size_t)slots * 8) != cudaSuccess || e->m_dropped.ensure(16) != cudaSuccess) {
            int r = fail(nullptr, CLK_ENOMEM, "cudaMalloc metrics table (%u slots)", slots); delete e; return r;
        }
        cudaMemset(e->m_keys.p, 0, (size_t)slots * 8); cudaMemset(e->m_counts.p, 0, (size_t)slots * 8); cudaMemset(e->m_dropped.p, 0, 16);
        e->T.m_mask = slots - 1;
    }
    *out = e;
    return CLK_OK;
} CLK_GUARD_INT(nullptr)

void clk_engine_destroy(clk_engine *e)
try {
    if (!e) return;
    if (!e->host_only) {
        cudaSetDevice(e->device);
        cudaDeviceSynchronize();
        e->snap[0].release(); e->snap[1].release();
        for (DevBuf *b : { &e->m_keys, &e->m_counts, &e->m_dropped, &e->g_keys, &e->g_counts, &e->g_dropped, &e->scan_sums, &e->scan_pos, &e->scan_total, &e->upd_out, &e->ctr_block }) b->release();
        for (int i = 0; i < 3; i++) e->work_rare[i].release();
        for (auto &kv : e->scratch) kv.second.release();
        for (auto &kv : e->varlen) {
            kv.second.buf.release();
            for (int a = 0; a < 3; a++) if (kv.second.ctx.aux[a]) cudaStreamDestroy(kv.second.ctx.aux[a]);
            for (int p = 0; p < 2; p++) { if (kv.second.ctx.ev_head[p]) cudaEventDestroy(kv.second.ctx.ev_head[p]); for (int a = 0; a < 3; a++) if (kv.second.ctx.ev_done[p][a]) cudaEventDestroy(kv.second.ctx.ev_done[p][a]); }
        }
        comm_release(e);
        for (int i = 0; i < 3; i++) { e->work_pkts[i].release(); e->work_offs[i].release(); e->work_verd[i].release(); if (e->pipe[i]) cudaStreamDestroy(e->pipe[i]); }
        if (e->ev0) cudaEventDestroy(e->ev0);
        if (e->ev1) cudaEventDestroy(e->ev1);
        if (e->stream) cudaStreamDestroy(e->stream);
    }
    delete e;
} CLK_GUARD_VOID

// ---- map ops ------------------------------------------------------------------------------------
int clk_map_update(clk_engine *e, int id, const void *key, size_t klen, const void *val, size_t vlen)
try {
    if (!e || !key || id < 0 || id >= CLK_MAP__COUNT) return fail(e, CLK_EINVAL, "bad map id %d", id);
    if (id == CLK_MAP_METRICS) return fail(e, CLK_EINVAL, "metrics_map is written by the classifier only");
    if (klen != kSpecs[id].klen || (id != CLK_MAP_CIDR && (vlen != kSpecs[id].vlen || !val))) return fail(e, CLK_EINVAL, "map %d: key/value size %zu/%zu, want %zu/%zu", id, klen, vlen, kSpecs[id].klen, kSpecs[id].vlen);
    std::lock_guard<std::mutex> g(e->mu);
    switch (id) {
    case CLK_MAP_CONTAINER: { uint64_t k; memcpy(&k, key, 8); if (!e->container.count(k) && e->container.size() >= e->max_containers) return fail(e, CLK_E2BIG, "container_map full (%u)", e->max_containers); memcpy(&e->container[k], val, sizeof(clk_container_config)); break; }
    case CLK_MAP_BYPASS: { uint64_t k; memcpy(&k, key, 8); if (!e->bypass.count(k) && e->bypass.size() >= e->max_containers) return fail(e, CLK_E2BIG, "bypass_map full"); e->bypass[k] = *(const uint8_t *)val; break; }
    case CLK_MAP_DNS_CACHE: { uint32_t k; memcpy(&k, key, 4); if (!e->dns.count(k) && e->dns.size() >= e->max_dns) return fail(e, CLK_E2BIG, "dns_cache full (%u)", e->max_dns); memcpy(&e->dns[k], val, sizeof(clk_dns_entry)); break; }
    case CLK_MAP_ROUTE: {
        clk_route_key rk; memcpy(&rk, key, sizeof rk); clk_route_val rv; memcpy(&rv, val, sizeof rv);
        uint64_t k = route_k(&rk);
        if (((k << 16) | rv.envoy_port) == SLOT_EMPTY) return fail(e, CLK_EINVAL, "route {0xffffffff,65535}->65535 is not representable");
        if (!e->route.count(k) && e->route.size() >= e->max_routes) return fail(e, CLK_E2BIG, "route_map full (%u)", e->max_routes);
        e->route[k] = rv.envoy_port; break;
    }
    case CLK_MAP_ENDPOINT: { uint32_t k; memcpy(&k, key, 4); memcpy(&e->endpoint[k], val, sizeof(clk_endpoint)); break; }
    case CLK_MAP_ENDPOINT6: { std::array<uint8_t, 16> k; memcpy(k.data(), key, 16); if (!e->endpoint6.count(k) && e->endpoint6.size() >= 256) return fail(e, CLK_E2BIG, "endpoint6 table full (256)"); memcpy(&e->endpoint6[k], val, sizeof(clk_endpoint)); break; }
    case CLK_MAP_CIDR: {
        clk_cidr c; memcpy(&c, key, sizeof c);
        if (!is_cidr_mask(c.mask)) return fail(e, CLK_EINVAL, "mask %#x is not a CIDR prefix mask", c.mask);
        if (!e->cidr.count({ c.addr, c.mask }) && e->cidr.size() >= e->max_cidrs) return fail(e, CLK_E2BIG, "cidr set full (%u)", e->max_cidrs);
        e->cidr[{ c.addr, c.mask }] = 1; break;
    }
    }
    e->dirty = true;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_map_delete(clk_engine *e, int id, const void *key, size_t klen)
try {
    if (!e || !key || id < 0 || id >= CLK_MAP__COUNT || id == CLK_MAP_METRICS) return fail(e, CLK_EINVAL, "bad map id %d", id);
    if (klen != kSpecs[id].klen) return fail(e, CLK_EINVAL, "map %d: key size %zu, want %zu", id, klen, kSpecs[id].klen);
    std::lock_guard<std::mutex> g(e->mu);
    size_t n = 0;
    switch (id) {
    case CLK_MAP_CONTAINER: { uint64_t k; memcpy(&k, key, 8); n = e->container.erase(k); break; }
    case CLK_MAP_BYPASS:    { uint64_t k; memcpy(&k, key, 8); n = e->bypass.erase(k); break; }
    case CLK_MAP_DNS_CACHE: { uint32_t k; memcpy(&k, key, 4); n = e->dns.erase(k); break; }
    case CLK_MAP_ROUTE:     { clk_route_key rk; memcpy(&rk, key, sizeof rk); n = e->route.erase(route_k(&rk)); break; }
    case CLK_MAP_ENDPOINT:  { uint32_t k; memcpy(&k, key, 4); n = e->endpoint.erase(k); break; }
    case CLK_MAP_ENDPOINT6: { std::array<uint8_t, 16> k; memcpy(k.data(), key, 16); n = e->endpoint6.erase(k); break; }
    case CLK_MAP_CIDR:      { clk_cidr c; memcpy(&c, key, sizeof c); n = e->cidr.erase({ c.addr, c.mask }); break; }
    }
    if (!n) return CLK_ENOENT;                          // ebpf.ErrKeyNotExist: callers treat as success (manager.go:222-231)
    e->dirty = true;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_map_lookup(clk_engine *e, int id, const void *key, size_t klen, void *val, size_t vlen)
try {
    if (!e || !key || !val || id < 0 || id >= CLK_MAP__COUNT || id == CLK_MAP_METRICS) return fail(e, CLK_EINVAL, "bad map id %d", id);
    if (klen != kSpecs[id].klen || vlen != kSpecs[id].vlen) return fail(e, CLK_EINVAL, "map %d: key/value size", id);
    std::lock_guard<std::mutex> g(e->mu);
    switch (id) {
    case CLK_MAP_CONTAINER: { uint64_t k; memcpy(&k, key, 8); auto it = e->container.find(k); if (it == e->container.end()) return CLK_ENOENT; memcpy(val, &it->second, vlen); return CLK_OK; }
    case CLK_MAP_BYPASS:    { uint64_t k; memcpy(&k, key, 8); auto it = e->bypass.find(k); if (it == e->bypass.end()) return CLK_ENOENT; memcpy(val, &it->second, 1); return CLK_OK; }
    case CLK_MAP_DNS_CACHE: { uint32_t k; memcpy(&k, key, 4); auto it = e->dns.find(k); if (it == e->dns.end()) return CLK_ENOENT; memcpy(val, &it->second, vlen); return CLK_OK; }
    case CLK_MAP_ROUTE:     { clk_route_key rk; memcpy(&rk, key, sizeof rk); auto it = e->route.find(route_k(&rk)); if (it == e->route.end()) return CLK_ENOENT; clk_route_val rv = { it->second, 0 }; memcpy(val, &rv, vlen); return CLK_OK; }
    case CLK_MAP_ENDPOINT:  { uint32_t k; memcpy(&k, key, 4); auto it = e->endpoint.find(k); if (it == e->endpoint.end()) return CLK_ENOENT; memcpy(val, &it->second, vlen); return CLK_OK; }
    case CLK_MAP_ENDPOINT6: { std::array<uint8_t, 16> k; memcpy(k.data(), key, 16); auto it = e->endpoint6.find(k); if (it == e->endpoint6.end()) return CLK_ENOENT; memcpy(val, &it->second, vlen); return CLK_OK; }
    case CLK_MAP_CIDR:      { clk_cidr c; memcpy(&c, key, sizeof c); auto it = e->cidr.find({ c.addr, c.mask }); if (it == e->cidr.end()) return CLK_ENOENT; *(uint8_t *)val = 1; return CLK_OK; }
    }
    return CLK_EINVAL;
} CLK_GUARD_INT(e)

int clk_map_count(clk_engine *e, int id, uint32_t *n)
try {
    if (!e || !n || id < 0 || id >= CLK_MAP__COUNT) return fail(e, CLK_EINVAL, "bad map id %d", id);
    std::lock_guard<std::mutex> g(e->mu);
    switch (id) {
    case CLK_MAP_CONTAINER: *n = (uint32_t)e->container.size(); break;
    case CLK_MAP_BYPASS: *n = (uint32_t)e->bypass.size(); break;
    case CLK_MAP_DNS_CACHE: *n = (uint32_t)e->dns.size(); break;
    case CLK_MAP_ROUTE: *n = (uint32_t)e->route.size(); break;
    case CLK_MAP_ENDPOINT: *n = (uint32_t)e->endpoint.size(); break;
    case CLK_MAP_ENDPOINT6: *n = (uint32_t)e->endpoint6.size(); break;
    case CLK_MAP_CIDR: *n = (uint32_t)e->cidr.size(); break;
    default: return fail(e, CLK_EINVAL, "use clk_counters_read for metrics_map");
    }
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_map_iter(clk_engine *e, int id, void *keys, void *vals, uint32_t cap, uint32_t *n)
try {
    if (!e || !n || id < 0 || id >= CLK_MAP__COUNT || id == CLK_MAP_METRICS) return fail(e, CLK_EINVAL, "bad map id %d", id);
    std::lock_guard<std::mutex> g(e->mu);
    uint8_t *kp = (uint8_t *)keys, *vp = (uint8_t *)vals; uint32_t i = 0;
    auto put = [&](const void *k, const void *v) { if (i < cap) { if (kp) memcpy(kp + (size_t)i * kSpecs[id].klen, k, kSpecs[id].klen); if (vp) memcpy(vp + (size_t)i * kSpecs[id].vlen, v, kSpecs[id].vlen); } i++; };
    switch (id) {
    case CLK_MAP_CONTAINER: for (auto &kv : e->container) put(&kv.first, &kv.second); break;
    case CLK_MAP_BYPASS:    for (auto &kv : e->bypass) put(&kv.first, &kv.second); break;
    case CLK_MAP_DNS_CACHE: for (auto &kv : e->dns) put(&kv.first, &kv.second); break;
    case CLK_MAP_ROUTE:     for (auto &kv : e->route) { clk_route_key rk = { (uint32_t)(kv.first >> 16), (uint16_t)(kv.first & 0xffff), 0 }; clk_route_val rv = { kv.second, 0 }; put(&rk, &rv); } break;
    case CLK_MAP_ENDPOINT:  for (auto &kv : e->endpoint) put(&kv.first, &kv.second); break;
    case CLK_MAP_ENDPOINT6: for (auto &kv : e->endpoint6) put(kv.first.data(), &kv.second); break;
    case CLK_MAP_CIDR:      for (auto &kv : e->cidr) { clk_cidr c = { kv.first.first, kv.first.second }; put(&c, &kv.second); } break;
    }
    *n = i;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_routes_replace(clk_engine *e, const clk_route *routes, uint32_t n)
try {
    if (!e || (n && !routes)) return fail(e, CLK_EINVAL, "routes is NULL");
    std::lock_guard<std::mutex> g(e->mu);
    if (n > e->max_routes) return fail(e, CLK_E2BIG, "%u routes > max_entries %u", n, e->max_routes);
    e->route.clear();
    for (uint32_t i = 0; i < n; i++) e->route[((uint64_t)routes[i].domain_hash << 16) | routes[i].dst_port] = routes[i].envoy_port;   // later duplicates win (UpdateAny)
    e->dirty = true;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_dns_update(clk_engine *e, uint32_t ip, uint32_t domain_hash_, uint32_t ttl, uint32_t now)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    if (!now) now = (uint32_t)time(nullptr);
    clk_dns_entry en = { domain_hash_, now + ttl };
    return clk_map_update(e, CLK_MAP_DNS_CACHE, &ip, 4, &en, sizeof en);
} CLK_GUARD_INT(e)

int clk_dns_gc(clk_engine *e, uint32_t now)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    if (!now) now = (uint32_t)time(nullptr);
    std::lock_guard<std::mutex> g(e->mu);
    int cleared = 0;
    for (auto it = e->dns.begin(); it != e->dns.end();) {
        if (it->second.expire_ts < now) { it = e->dns.erase(it); cleared++; } else ++it;       // manager.go:449-453
    }
    if (cleared) e->dirty = true;
    return cleared;
} CLK_GUARD_INT(e)

// ---- rules ----------------------------------------------------------------------------------------
int clk_validate_dst(const char *dst) try { return (dst && validate_dst(dst)) ? CLK_OK : CLK_EINVAL; } CLK_GUARD_INT(nullptr)
uint32_t clk_domain_hash(const char *domain, size_t len) try { return domain_hash(domain ? domain : "", domain ? len : 0); } CLK_GUARD_VAL(0u)
int clk_ip_to_u32(const char *s, uint32_t *out) try { uint32_t v = 0; bool ok = s && ip_to_u32(s, &v); if (out) *out = v; return ok ? CLK_OK : CLK_EINVAL; } CLK_GUARD_INT(nullptr)
int clk_cidr_to_addr_mask(const char *s, uint32_t *addr, uint32_t *mask)
try {
    uint32_t a = 0, m = 0; bool ok = s && cidr_to_addr_mask(s, &a, &m);
    if (addr) *addr = a; if (mask) *mask = m;
    return ok ? CLK_OK : CLK_EINVAL;
} CLK_GUARD_INT(nullptr)

int clk_rules_compile(clk_engine *e, const clk_rule *rules, uint32_t n, const clk_ports *ports, uint32_t *n_warnings)
try {
    if (!e || (n && !rules)) return fail(e, CLK_EINVAL, "rules is NULL");
    std::vector<EgressRule> rs(n);
    for (uint32_t i = 0; i < n; i++) {
        rs[i].dst = rules[i].dst ? rules[i].dst : ""; rs[i].proto = rules[i].proto ? rules[i].proto : "";
        rs[i].action = rules[i].action ? rules[i].action : ""; rs[i].port = rules[i].port;
        rs[i].path_default = rules[i].path_default ? rules[i].path_default : "";
        if (rules[i].n_path_rules && !rules[i].path_rules) return fail(e, CLK_EINVAL, "rule %u: path_rules is NULL", i);
        for (uint32_t k = 0; k < rules[i].n_path_rules; k++)
            rs[i].path_rules.push_back({ rules[i].path_rules[k].path ? rules[i].path_rules[k].path : "", rules[i].path_rules[k].action ? rules[i].path_rules[k].action : "" });
    }
    int base = ports && ports->tcp_port_base ? ports->tcp_port_base : 10001;
    CompiledRules c = compile_rules(rs, base);
    std::lock_guard<std::mutex> g(e->mu);
    if (c.mappings.size() > e->max_routes) return fail(e, CLK_E2BIG, "%zu TCP mappings > route_map max_entries %u", c.mappings.size(), e->max_routes);
    if (c.cidrs.size() > e->max_cidrs) return fail(e, CLK_E2BIG, "%zu CIDR rules > %u", c.cidrs.size(), e->max_cidrs);
    e->route.clear();                                    // syncRoutes -> SyncRoutes: clear then Update each (manager.go:347-386)
    for (auto &m : c.mappings) e->route[((uint64_t)domain_hash(m.dst.data(), m.dst.size()) << 16) | (uint16_t)m.dst_port] = (uint16_t)m.envoy_port;
    e->cidr.clear();
    for (auto &cm : c.cidrs) e->cidr[{ cm.first, cm.second }] = 1;
    if (n_warnings) *n_warnings = c.warnings;
    e->compiled = std::move(c);
    e->dirty = true;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_rules_dump(clk_engine *e, int kind, char *buf, size_t cap)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    std::lock_guard<std::mutex> g(e->mu);
    std::string s;
    if (kind == 0) for (auto &z : e->compiled.zones) s += z + "\n";
    else if (kind == 1) for (auto &z : e->compiled.snis) s += z + "\n";
    else if (kind == 2) for (auto &m : e->compiled.mappings) s += m.dst + ":" + std::to_string(m.dst_port) + ":" + std::to_string(m.envoy_port) + "\n";
    else if (kind == 3) for (auto &v : e->compiled.http) {               // one virtual host per line: domains | routes (A = allow, D = deny)
        for (size_t i = 0; i < v.domains.size(); i++) s += (i ? "," : "") + v.domains[i];
        s += "|";
        for (size_t i = 0; i < v.routes.size(); i++) s += std::string(i ? "," : "") + (v.routes[i].allow ? "A" : "D") + v.routes[i].prefix;
        s += "\n";
    }
    else return fail(e, CLK_EINVAL, "kind %d", kind);
    if (buf && cap) { size_t k = std::min(cap - 1, s.size()); memcpy(buf, s.data(), k); buf[k] = 0; }
    return (int)s.size() + 1;
} CLK_GUARD_INT(e)

int clk_publish(clk_engine *e)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    std::lock_guard<std::mutex> g(e->mu);
    return publish_locked(e);
} CLK_GUARD_INT(e)

// ---- classification ---------------------------------------------------------------------------------
static int finish_launch(clk_engine *e, const char *err, const LaunchInfo &li)
{
    if (err) return fail(e, CLK_EIO, "kernel launch failed: %s", err);
    e->stats.kernel_launches += li.launches ? li.launches : 1; e->stats.last_grid = li.grid; e->stats.last_block = li.block; e->stats.last_smem_bytes = li.smem;
    return CLK_OK;
}

int clk_classify_slots_dev(clk_engine *e, const void *d_pkts, uint32_t stride, uint64_t n, uint32_t flags, void *d_verdicts, void *stream)
try {
    if (!e || (n && (!d_pkts || !d_verdicts))) return fail(e, CLK_EINVAL, "NULL buffer");
    if (stride < 64 || stride % 16) return fail(e, CLK_EINVAL, "stride %u: must be >= 64 and a multiple of 16", stride);
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    LaunchInfo li = { 0, 0, 0, 0 };
    if (!stream) cudaEventRecord(e->ev0, st);                 // a caller that brings its own stream brings its own events
    DevBuf &sc = e->scratch[st];                              // per stream: launches on one stream are ordered, others get their own
    CUDA_TRY(e, sc.ensure((n + (1u << 20)) * 4));              // per-CTA regions (packets of the CTA's tiles); grows only for larger batches
    const char *err = launch_slots(e->T, d_pkts, stride, n, flags, (uint32_t *)d_verdicts, st, e->loader, e->stages,
                                   (uint32_t *)sc.p + 2, (uint32_t *)sc.p, &li);
    if (!stream) { cudaEventRecord(e->ev1, st); e->ev_recorded = true; }
    return finish_launch(e, err, li);
} CLK_GUARD_INT(e)

int clk_classify_varlen_dev(clk_engine *e, const void *d_pkts, const void *d_offs16, uint64_t n, uint32_t flags, void *d_verdicts, void *stream)
try {
    if (!e || (n && (!d_pkts || !d_verdicts || !d_offs16))) return fail(e, CLK_EINVAL, "NULL buffer");
    if (((uintptr_t)d_pkts & 15u) || ((uintptr_t)d_offs16 & 3u) || ((uintptr_t)d_verdicts & 3u)) return fail(e, CLK_EINVAL, "d_pkts must be 16-byte aligned (frames are read with 16-byte loads), offs16 / verdicts 4-byte aligned");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    LaunchInfo li = { 0, 0, 0, 0 };
    if (!stream) cudaEventRecord(e->ev0, st);                 // a caller that brings its own stream brings its own events
    const VarlenCtx *vc = nullptr;
    rc = varlen_ctx(e, st, n, &vc); if (rc) return rc;
    const char *err = launch_varlen(e->T, d_pkts, (const uint32_t *)d_offs16, n, flags, (uint32_t *)d_verdicts, st, vc, &li);
    if (!stream) { cudaEventRecord(e->ev1, st); e->ev_recorded = true; }
    return finish_launch(e, err, li);
} CLK_GUARD_INT(e)

// Host-buffer entry: chunked 3-deep pipeline, each chunk = H2D copy -> kernel -> D2H copy on its own stream.
static int classify_host(clk_engine *e, const uint8_t *pkts, const uint32_t *offs16, uint32_t stride, uint64_t n, uint32_t flags, uint32_t *verdicts)
{
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    CUDA_TRY(e, cudaSetDevice(e->device));
    const uint64_t chunk = e->chunk_pkts;
    uint64_t done = 0; int k = 0;
    while (done < n) {
        uint64_t cnt = std::min<uint64_t>(chunk, n - done);
        if (offs16) {                                      // keep a chunk under 256 MiB of payload
            const uint64_t cap16 = (256ull << 20) / 16;
            while (cnt > 1 && (uint64_t)(offs16[done + cnt] - offs16[done]) > cap16) cnt /= 2;
            // the chunk's byte range must lie inside the caller's blob before anything is copied from it
            if (offs16[done + cnt] < offs16[done] || offs16[done + cnt] > offs16[n]) {
                for (int q = 0; q < 3; q++) cudaStreamSynchronize(e->pipe[q]);      // earlier chunks still copy into the caller's verdicts: finish them before returning
                return fail(e, CLK_EINVAL, "offs16 not monotonic in [%llu, %llu]", (unsigned long long)done, (unsigned long long)(done + cnt));
            }
        }
        int s = k % 3; cudaStream_t st = e->pipe[s];
        CUDA_TRY(e, cudaStreamSynchronize(st));            // buffers of this lane are free again
        LaunchInfo li = { 0, 0, 0, 0 }; const char *err;
        CUDA_TRY(e, e->work_verd[s].ensure(cnt * 4));
        if (offs16) {
            uint64_t b0 = (uint64_t)offs16[done] * 16, b1 = (uint64_t)offs16[done + cnt] * 16;
            CUDA_TRY(e, e->work_pkts[s].ensure(b1 - b0 + 16));
            CUDA_TRY(e, e->work_offs[s].ensure((cnt + 1) * 4));
            // offsets first: if they sit in pageable memory the (blocking, staged) copy is short and does not wait behind the frames
            CUDA_TRY(e, cudaMemcpyAsync(e->work_offs[s].p, offs16 + done, (cnt + 1) * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(e, cudaMemcpyAsync(e->work_pkts[s].p, pkts + b0, b1 - b0, cudaMemcpyHostToDevice, st));
            // offsets must be monotonic (the kernel trusts them): checked while this chunk's copies are on the wire, before its launch
            {
                uint32_t bad = 0;
                const uint32_t *o = offs16 + done;
                for (uint64_t i = 0; i < cnt; i++) bad |= (uint32_t)(o[i + 1] < o[i]);
                if (bad) {
                    for (int q = 0; q < 3; q++) cudaStreamSynchronize(e->pipe[q]);      // nothing of ours touches the caller's buffers after we return
                    for (uint64_t i = 0; i < cnt; i++) if (o[i + 1] < o[i]) return fail(e, CLK_EINVAL, "offs16 not monotonic at %llu", (unsigned long long)(done + i));
                }
            }
            // offsets stay absolute: bias the base pointer instead of rewriting them
            const VarlenCtx *vc = nullptr;
            rc = varlen_ctx(e, st, cnt, &vc); if (rc) return rc;
            err = launch_varlen(e->T, (const uint8_t *)e->work_pkts[s].p - b0, (const uint32_t *)e->work_offs[s].p, cnt, flags, (uint32_t *)e->work_verd[s].p, st, vc, &li);
        } else {
            CUDA_TRY(e, e->work_pkts[s].ensure(cnt * stride));
            CUDA_TRY(e, cudaMemcpyAsync(e->work_pkts[s].p, pkts + done * stride, cnt * stride, cudaMemcpyHostToDevice, st));
            CUDA_TRY(e, e->work_rare[s].ensure((cnt + (1u << 20)) * 4));
            err = launch_slots(e->T, e->work_pkts[s].p, stride, cnt, flags, (uint32_t *)e->work_verd[s].p, st, e->loader, e->stages,
                               (uint32_t *)e->work_rare[s].p + 2, (uint32_t *)e->work_rare[s].p, &li);
        }
        rc = finish_launch(e, err, li); if (rc) return rc;
        CUDA_TRY(e, cudaMemcpyAsync(verdicts + done, e->work_verd[s].p, cnt * 4, cudaMemcpyDeviceToHost, st));
        done += cnt; k++;
    }
    for (int s = 0; s < 3; s++) CUDA_TRY(e, cudaStreamSynchronize(e->pipe[s]));
    return CLK_OK;
}

int clk_classify_slots(clk_engine *e, const uint8_t *pkts, uint32_t stride, uint64_t n, uint32_t flags, uint32_t *verdicts)
try {
    if (!e || (n && (!pkts || !verdicts))) return fail(e, CLK_EINVAL, "NULL buffer");
    if (stride < 64 || stride % 16) return fail(e, CLK_EINVAL, "stride %u: must be >= 64 and a multiple of 16", stride);
    return classify_host(e, pkts, nullptr, stride, n, flags, verdicts);
} CLK_GUARD_INT(e)

int clk_classify_varlen(clk_engine *e, const uint8_t *pkts, const uint32_t *offs16, uint64_t n, uint32_t flags, uint32_t *verdicts)
try {
    if (!e || (n && (!pkts || !verdicts || !offs16))) return fail(e, CLK_EINVAL, "NULL buffer");
    return classify_host(e, pkts, offs16, 0, n, flags, verdicts);
} CLK_GUARD_INT(e)

// ---- compact ingest ------------------------------------------------------------------------------------------------
int clk_pack_hdr16(const uint8_t *pkts, uint32_t stride, uint64_t n, clk_hdr16 *out, uint8_t *aux, uint64_t aux_cap, uint64_t *n_aux)
try {
    if ((n && (!pkts || !out)) || !n_aux || stride < 64) return fail(nullptr, CLK_EINVAL, "NULL buffer or stride < 64");
    uint64_t na = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *p = pkts + i * stride;
        clk_hdr16 r;
        if (clk_hdr16_from_frame(p, r)) {
            if (!aux || na >= aux_cap) return fail(nullptr, CLK_E2BIG, "aux holds %llu slots, more frames need one", (unsigned long long)aux_cap);
            if (na > 0xffffffffull) return fail(nullptr, CLK_E2BIG, "more than 2^32 aux slots");
            memcpy(aux + na * 64, p, 64);
            r.kind = CLK_HDR_AUX; r.src = (uint32_t)na++;
        }
        out[i] = r;
    }
    *n_aux = na;
    return CLK_OK;
} CLK_GUARD_INT(nullptr)

int clk_classify_hdr16_dev(clk_engine *e, const void *d_recs, uint64_t n, const void *d_aux, uint32_t flags, void *d_verdicts, void *stream)
try {
    if (!e || (n && (!d_recs || !d_verdicts))) return fail(e, CLK_EINVAL, "NULL buffer");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    LaunchInfo li = { 0, 0, 0, 0 };
    const char *err = launch_hdr16(e->T, d_recs, d_aux, n, flags, (uint32_t *)d_verdicts, stream ? (cudaStream_t)stream : e->stream, &li);
    return finish_launch(e, err, li);
} CLK_GUARD_INT(e)

int clk_classify_hdr16(clk_engine *e, const clk_hdr16 *recs, uint64_t n, const uint8_t *aux, uint64_t n_aux, uint32_t flags, uint32_t *verdicts)
try {
    if (!e || (n && (!recs || !verdicts)) || (n_aux && !aux)) return fail(e, CLK_EINVAL, "NULL buffer");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    CUDA_TRY(e, cudaSetDevice(e->device));
    for (int s = 0; s < 3; s++) CUDA_TRY(e, cudaStreamSynchronize(e->pipe[s]));
    if (n_aux) {                                              // the few whole frames first, once
        CUDA_TRY(e, e->work_offs[0].ensure(n_aux * 64));
        CUDA_TRY(e, cudaMemcpyAsync(e->work_offs[0].p, aux, n_aux * 64, cudaMemcpyHostToDevice, e->pipe[0]));
        CUDA_TRY(e, cudaStreamSynchronize(e->pipe[0]));
    }
    const uint64_t chunk = e->chunk_pkts * 2;
    uint64_t done = 0; int k = 0;
    while (done < n) {
        const uint64_t cnt = std::min<uint64_t>(chunk, n - done);
        const int s = k % 3; cudaStream_t st = e->pipe[s];
        CUDA_TRY(e, cudaStreamSynchronize(st));
        CUDA_TRY(e, e->work_pkts[s].ensure(cnt * sizeof(clk_hdr16))); CUDA_TRY(e, e->work_verd[s].ensure(cnt * 4));
        CUDA_TRY(e, cudaMemcpyAsync(e->work_pkts[s].p, recs + done, cnt * sizeof(clk_hdr16), cudaMemcpyHostToDevice, st));
        LaunchInfo li = { 0, 0, 0, 0 };
        const char *err = launch_hdr16(e->T, e->work_pkts[s].p, n_aux ? e->work_offs[0].p : nullptr, cnt, flags, (uint32_t *)e->work_verd[s].p, st, &li);
        rc = finish_launch(e, err, li); if (rc) return rc;
        CUDA_TRY(e, cudaMemcpyAsync(verdicts + done, e->work_verd[s].p, cnt * 4, cudaMemcpyDeviceToHost, st));
        done += cnt; k++;
    }
    for (int s = 0; s < 3; s++) CUDA_TRY(e, cudaStreamSynchronize(e->pipe[s]));
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_classify_events(clk_engine *e, const clk_event *events, uint64_t n, uint32_t flags, uint32_t *verdicts)
try {
    if (!e || (n && (!events || !verdicts))) return fail(e, CLK_EINVAL, "NULL buffer");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    if (!n) return CLK_OK;
    CUDA_TRY(e, cudaSetDevice(e->device));
    cudaStream_t st = e->pipe[0];
    CUDA_TRY(e, e->work_pkts[0].ensure(n * sizeof(clk_event)));
    CUDA_TRY(e, e->work_verd[0].ensure(n * 4));
    CUDA_TRY(e, cudaMemcpyAsync(e->work_pkts[0].p, events, n * sizeof(clk_event), cudaMemcpyHostToDevice, st));
    LaunchInfo li = { 0, 0, 0, 0 };
    const char *err = launch_events(e->T, (const clk_event *)e->work_pkts[0].p, n, flags, (uint32_t *)e->work_verd[0].p, st, &li);
    rc = finish_launch(e, err, li); if (rc) return rc;
    CUDA_TRY(e, cudaMemcpyAsync(verdicts, e->work_verd[0].p, n * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_dns_answers(clk_engine *e, const uint8_t *pkts, const uint32_t *offs16, uint64_t n, const uint32_t *verdicts,
                    clk_dns_update_rec *out, uint64_t cap, uint64_t *n_out, int apply, uint32_t now)
try {
    if (!e || !n_out || (n && (!pkts || !offs16 || !verdicts))) return fail(e, CLK_EINVAL, "NULL buffer");
    NEED_GPU(e);
    *n_out = 0;
    if (!n) return CLK_OK;
    std::vector<clk_dns_update_rec> host_out;
    {
        std::lock_guard<std::mutex> g(e->mu);
        int rc = ensure_published(e); if (rc) return rc;
        CUDA_TRY(e, cudaSetDevice(e->device));
        cudaStream_t st = e->pipe[0];
        uint64_t bytes = (uint64_t)offs16[n] * 16, total = 0;
        for (uint64_t i = 0; i < n; i++) if (CLK_HOOK(verdicts[i]) == CLK_HOOK_NONE) total += CLK_NUPD(verdicts[i]);
        uint32_t nb = (uint32_t)((n + 4095) / 4096);
        CUDA_TRY(e, e->work_pkts[0].ensure(bytes + 16)); CUDA_TRY(e, e->work_offs[0].ensure((n + 1) * 4)); CUDA_TRY(e, e->work_verd[0].ensure(n * 4));
        CUDA_TRY(e, e->scan_sums.ensure((size_t)nb * 8)); CUDA_TRY(e, e->scan_pos.ensure(n * 8)); CUDA_TRY(e, e->scan_total.ensure(8));
        CUDA_TRY(e, e->upd_out.ensure((total + 1) * sizeof(clk_dns_update_rec)));
        CUDA_TRY(e, cudaMemcpyAsync(e->work_pkts[0].p, pkts, bytes, cudaMemcpyHostToDevice, st));
        CUDA_TRY(e, cudaMemcpyAsync(e->work_offs[0].p, offs16, (n + 1) * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(e, cudaMemcpyAsync(e->work_verd[0].p, verdicts, n * 4, cudaMemcpyHostToDevice, st));
        const char *err = launch_dns_extract(e->T, e->work_pkts[0].p, (const uint32_t *)e->work_offs[0].p, n, (const uint32_t *)e->work_verd[0].p,
                                             (unsigned long long *)e->scan_sums.p, (unsigned long long *)e->scan_pos.p, (unsigned long long *)e->scan_total.p,
                                             (clk_dns_update_rec *)e->upd_out.p, total, st);
        if (err) return fail(e, CLK_EIO, "dns extract launch failed: %s", err);
        e->stats.kernel_launches += 4;
        unsigned long long dev_total = 0;
        CUDA_TRY(e, cudaMemcpyAsync(&dev_total, e->scan_total.p, 8, cudaMemcpyDeviceToHost, st));
        host_out.resize(total);
        if (total) CUDA_TRY(e, cudaMemcpyAsync(host_out.data(), e->upd_out.p, total * sizeof(clk_dns_update_rec), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(e, cudaStreamSynchronize(st));
        if (dev_total != total) return fail(e, CLK_EIO, "dns extract: device total %llu != host total %llu", dev_total, (unsigned long long)total);
    }
    *n_out = host_out.size();
    if (out) memcpy(out, host_out.data(), std::min<uint64_t>(cap, host_out.size()) * sizeof(clk_dns_update_rec));
    if (apply) for (auto &u : host_out) { int rc = clk_dns_update(e, u.ip, u.domain_hash, u.ttl, now); if (rc && rc != CLK_E2BIG) return rc; }
    return CLK_OK;
} CLK_GUARD_INT(e)

// ---- counters -----------------------------------------------------------------------------------------
int clk_counters_block_rows(clk_engine *e, uint32_t *rows)
try {
    if (!e || !rows) return fail(e, CLK_EINVAL, "NULL");
    *rows = e->max_metrics + 1 + 1024;     // header + max_entries (+ slack: the entry cap is enforced without a lock)
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_counters_export(clk_engine *e, void *d_block, void *stream)
try {
    if (!e || !d_block) return fail(e, CLK_EINVAL, "NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    const char *err = launch_counters_export(e->T, (clk_metric *)d_block, e->max_metrics + 1 + 1024, stream ? (cudaStream_t)stream : e->stream);
    if (err) return fail(e, CLK_EIO, "counters export: %s", err);
    e->stats.kernel_launches++;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_totals_rows(clk_engine *e, uint32_t *rows)
try {
    if (!e || !rows) return fail(e, CLK_EINVAL, "NULL");
    *rows = 2 * e->max_containers;      // container_map and bypass_map are separate maps of max_containers rows each
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_totals_export(clk_engine *e, void *d_block, void *stream)
try {
    if (!e || !d_block) return fail(e, CLK_EINVAL, "NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    if (e->T.n_crows > 2 * e->max_containers)
        return fail(e, CLK_E2BIG, "%u cgroup rows do not fit the %u-row totals block (cgroups that left both maps but still own metric rows: clk_counters_reset reclaims them)", e->T.n_crows, 2 * e->max_containers);
    const char *err = launch_counters_totals(e->T, (clk_total *)d_block, 2 * e->max_containers, stream ? (cudaStream_t)stream : e->stream);
    if (err) return fail(e, CLK_EIO, "totals export: %s", err);
    e->stats.kernel_launches++;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_totals_merge(clk_engine *e, const void *d_blocks, uint32_t n_blocks, clk_total *out, uint32_t cap, uint32_t *n)
try {
    if (!e || !d_blocks || !n || (cap && !out)) return fail(e, CLK_EINVAL, "NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    std::vector<clk_total> rows((size_t)n_blocks * 2 * e->max_containers);
    if (!rows.empty()) CUDA_TRY(e, cudaMemcpy(rows.data(), d_blocks, rows.size() * sizeof(clk_total), cudaMemcpyDeviceToHost));
    std::map<uint64_t, clk_total> sum;
    for (const clk_total &r : rows) {
        if (r.cgroup_id == 0 && !(r.count[0] | r.count[1] | r.count[2])) continue;
        clk_total &t = sum[r.cgroup_id]; t.cgroup_id = r.cgroup_id;
        for (int a = 0; a < 3; a++) t.count[a] += r.count[a];
    }
    *n = (uint32_t)sum.size();
    uint32_t k = 0;
    for (auto &kv : sum) { if (k >= cap) break; out[k++] = kv.second; }
    return CLK_OK;
} CLK_GUARD_INT(e)

static DevTables global_view(clk_engine *e)
{
    DevTables G = e->T;
    G.m_keys = (unsigned long long *)e->g_keys.p; G.m_counts = (unsigned long long *)e->g_counts.p; G.m_dropped = (unsigned long long *)e->g_dropped.p; G.m_max = 0xffffffffu; G.m_mask = e->g_mask;
    return G;
}

int clk_counters_import(clk_engine *e, const void *d_blocks, uint32_t n_blocks, void *stream)
try {
    if (!e || !d_blocks) return fail(e, CLK_EINVAL, "NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    size_t slots = (size_t)e->T.m_mask + 1;                     // every rank may contribute max_entries distinct keys
    while (slots < 2 * (size_t)n_blocks * ((size_t)e->max_metrics + 1) && slots < ((size_t)1 << 31)) slots <<= 1;
    if (slots - 1 > e->g_mask) e->g_mask = (uint32_t)(slots - 1);
    slots = (size_t)e->g_mask + 1;
    CUDA_TRY(e, e->g_keys.ensure(slots * 8)); CUDA_TRY(e, e->g_counts.ensure(slots * 8)); CUDA_TRY(e, e->g_dropped.ensure(16));
    DevTables G = global_view(e);
    const char *err = launch_counters_clear(G, st);
    if (!err) err = launch_counters_import(G, (const clk_metric *)d_blocks, n_blocks, e->max_metrics + 1 + 1024, st);
    if (err) return fail(e, CLK_EIO, "counters import: %s", err);
    e->stats.kernel_launches += 2;
    return CLK_OK;
} CLK_GUARD_INT(e)

static int counters_read_impl(clk_engine *e, bool global, clk_metric *out, uint32_t cap, uint32_t *n)
{
    if (!e || !n) return fail(e, CLK_EINVAL, "NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    if (global && !e->g_keys.p) { *n = 0; return CLK_OK; }
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    DevTables V = global ? global_view(e) : e->T;
    uint32_t rows = V.m_mask + 2;                               // header + every slot (the merged view may hold more than max_entries)
    CUDA_TRY(e, e->ctr_block.ensure((size_t)rows * sizeof(clk_metric)));
    const char *err = launch_counters_export(V, (clk_metric *)e->ctr_block.p, rows, e->stream);
    if (err) return fail(e, CLK_EIO, "counters export: %s", err);
    e->stats.kernel_launches++;
    std::vector<clk_metric> host(rows);
    CUDA_TRY(e, cudaMemcpyAsync(host.data(), e->ctr_block.p, (size_t)rows * sizeof(clk_metric), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(&e->stats.metrics_dropped, V.m_dropped, 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    uint32_t k = 0; const uint64_t valid = std::min<uint64_t>(host[0].count, rows - 1);
    for (uint64_t i = 1; i <= valid; i++) if (host[i].count) { if (out && k < cap) out[k] = host[i]; k++; }
    *n = k;
    return CLK_OK;
}

int clk_counters_read(clk_engine *e, clk_metric *out, uint32_t cap, uint32_t *n) try { return counters_read_impl(e, false, out, cap, n); } CLK_GUARD_INT(e)
int clk_counters_read_global(clk_engine *e, clk_metric *out, uint32_t cap, uint32_t *n) try { return counters_read_impl(e, true, out, cap, n); } CLK_GUARD_INT(e)

int clk_counters_reset(clk_engine *e)
try {
    if (!e) return fail(e, CLK_EINVAL, "NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    const char *err = launch_counters_clear(e->T, e->stream);
    if (err) return fail(e, CLK_EIO, "counters clear: %s", err);
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    for (uint64_t cg : e->crow_cg) if (!e->container.count(cg) && !e->bypass.count(cg)) { e->dirty = true; break; }   // dead rows: the next publish drops them
    e->stats.kernel_launches++;
    return CLK_OK;
} CLK_GUARD_INT(e)

// ---- pinned state (the bpffs analogue) ------------------------------------------------
namespace {
const char kPinMagic[8] = { 'C', 'L', 'K', 'P', 'I', 'N', 1, 0 };
struct PinWriter {
    std::string b;
    void raw(const void *p, size_t n) { b.append((const char *)p, n); }
    void u32(uint32_t v) { raw(&v, 4); }
    void u64(uint64_t v) { raw(&v, 8); }
    void str(const std::string &s) { u32((uint32_t)s.size()); raw(s.data(), s.size()); }
    void section(uint32_t tag, uint64_t count) { u32(tag); u64(count); }
};
struct PinReader {
    const std::string &b; size_t p = 0; bool ok = true;
    explicit PinReader(const std::string &s) : b(s) {}
    bool raw(void *out, size_t n) { if (!ok || p + n > b.size()) { ok = false; memset(out, 0, n); return false; } memcpy(out, b.data() + p, n); p += n; return true; }
    uint32_t u32() { uint32_t v; raw(&v, 4); return v; }
    uint64_t u64() { uint64_t v; raw(&v, 8); return v; }
    std::string str() { uint32_t n = u32(); if (!ok || p + n > b.size()) { ok = false; return ""; } std::string s(b.data() + p, n); p += n; return s; }
};
enum { PIN_END = 0, PIN_CONTAINER, PIN_BYPASS, PIN_DNS, PIN_ROUTE, PIN_EP, PIN_EP6, PIN_CIDR, PIN_ZONES, PIN_SNIS, PIN_MAPPINGS, PIN_CCIDRS, PIN_HTTP };
}  // namespace

int clk_state_save(clk_engine *e, const char *path)
try {
    if (!e || !path || !*path) return fail(e, CLK_EINVAL, "NULL");
    PinWriter w;
    {
        std::lock_guard<std::mutex> g(e->mu);
        w.raw(kPinMagic, 8);
        w.section(PIN_CONTAINER, e->container.size()); for (auto &kv : e->container) { w.u64(kv.first); w.raw(&kv.second, sizeof kv.second); }
        w.section(PIN_BYPASS, e->bypass.size());       for (auto &kv : e->bypass) { w.u64(kv.first); w.raw(&kv.second, 1); }
        w.section(PIN_DNS, e->dns.size());             for (auto &kv : e->dns) { w.u32(kv.first); w.raw(&kv.second, sizeof kv.second); }
        w.section(PIN_ROUTE, e->route.size());         for (auto &kv : e->route) { w.u64(kv.first); w.raw(&kv.second, 2); }
        w.section(PIN_EP, e->endpoint.size());         for (auto &kv : e->endpoint) { w.u32(kv.first); w.raw(&kv.second, sizeof kv.second); }
        w.section(PIN_EP6, e->endpoint6.size());       for (auto &kv : e->endpoint6) { w.raw(kv.first.data(), 16); w.raw(&kv.second, sizeof kv.second); }
        w.section(PIN_CIDR, e->cidr.size());           for (auto &kv : e->cidr) { w.u32(kv.first.first); w.u32(kv.first.second); }
        w.section(PIN_ZONES, e->compiled.zones.size()); for (auto &z : e->compiled.zones) w.str(z);
        w.section(PIN_SNIS, e->compiled.snis.size());   for (auto &z : e->compiled.snis) w.str(z);
        w.section(PIN_MAPPINGS, e->compiled.mappings.size()); for (auto &m : e->compiled.mappings) { w.str(m.dst); w.u32((uint32_t)m.dst_port); w.u32((uint32_t)m.envoy_port); }
        w.section(PIN_CCIDRS, e->compiled.cidrs.size()); for (auto &c : e->compiled.cidrs) { w.u32(c.first); w.u32(c.second); }
        w.section(PIN_HTTP, e->compiled.http.size());
        for (auto &v : e->compiled.http) {
            w.u32((uint32_t)v.domains.size()); for (auto &d : v.domains) w.str(d);
            w.u32((uint32_t)v.routes.size());  for (auto &r : v.routes) { w.u32(r.allow ? 1u : 0u); w.str(r.prefix); }
        }
        w.section(PIN_END, 0);
    }
    const std::string tmp = std::string(path) + ".tmp." + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(e, CLK_EIO, "open %s: %s", tmp.c_str(), strerror(errno));
    const bool wrote = fwrite(w.b.data(), 1, w.b.size(), f) == w.b.size();
    const bool closed = fclose(f) == 0;
    if (!wrote || !closed || rename(tmp.c_str(), path) != 0) { int en = errno; remove(tmp.c_str()); return fail(e, CLK_EIO, "write %s: %s", path, strerror(en)); }
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_state_load(clk_engine *e, const char *path)
try {
    if (!e || !path || !*path) return fail(e, CLK_EINVAL, "NULL");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(e, errno == ENOENT ? CLK_ENOENT : CLK_EIO, "open %s: %s", path, strerror(errno));
    std::string buf; char chunk[1 << 16]; size_t k;
    while ((k = fread(chunk, 1, sizeof chunk, f)) > 0) buf.append(chunk, k);
    fclose(f);
    PinReader r(buf);
    char magic[8]; r.raw(magic, 8);
    if (!r.ok || memcmp(magic, kPinMagic, 8) != 0) return fail(e, CLK_EINVAL, "%s is not a clawker-gpu state file (or a different version)", path);
    // parse into fresh tables first: a truncated or oversized file must leave the engine untouched
    decltype(e->container) container; decltype(e->bypass) bypass; decltype(e->dns) dns; decltype(e->route) route;
    decltype(e->endpoint) endpoint; decltype(e->endpoint6) endpoint6; decltype(e->cidr) cidr; CompiledRules comp;
    for (;;) {
        const uint32_t tag = r.u32(); const uint64_t n = r.u64();
        if (!r.ok) return fail(e, CLK_EINVAL, "%s: truncated", path);
        if (tag == PIN_END) break;
        if (n > (buf.size() - r.p)) return fail(e, CLK_EINVAL, "%s: section %u claims %llu rows", path, tag, (unsigned long long)n);
        for (uint64_t i = 0; i < n && r.ok; i++) {
            switch (tag) {
            case PIN_CONTAINER: { uint64_t cg = r.u64(); clk_container_config c; r.raw(&c, sizeof c); container[cg] = c; break; }
            case PIN_BYPASS:    { uint64_t cg = r.u64(); uint8_t v; r.raw(&v, 1); bypass[cg] = v; break; }
            case PIN_DNS:       { uint32_t ip = r.u32(); clk_dns_entry d; r.raw(&d, sizeof d); dns[ip] = d; break; }
            case PIN_ROUTE:     { uint64_t rk = r.u64(); uint16_t v; r.raw(&v, 2); route[rk] = v; break; }
            case PIN_EP:        { uint32_t ip = r.u32(); clk_endpoint ep; r.raw(&ep, sizeof ep); endpoint[ip] = ep; break; }
            case PIN_EP6:       { std::array<uint8_t, 16> a; r.raw(a.data(), 16); clk_endpoint ep; r.raw(&ep, sizeof ep); endpoint6[a] = ep; break; }
            case PIN_CIDR:      { uint32_t a = r.u32(), m = r.u32(); if (!is_cidr_mask(m)) return fail(e, CLK_EINVAL, "%s: mask %#x is not a CIDR prefix mask", path, m); cidr[{ a, m }] = 1; break; }
            case PIN_ZONES:     comp.zones.push_back(r.str()); break;
            case PIN_SNIS:      comp.snis.push_back(r.str()); break;
            case PIN_MAPPINGS:  { TCPMapping m; m.dst = r.str(); m.dst_port = (int)r.u32(); m.envoy_port = (int)r.u32(); comp.mappings.push_back(m); break; }
            case PIN_CCIDRS:    { uint32_t a = r.u32(), m = r.u32(); comp.cidrs.push_back({ a, m }); break; }
            case PIN_HTTP: {
                HttpVhost v; uint32_t nd = r.u32();
                for (uint32_t k = 0; k < nd && r.ok; k++) v.domains.push_back(r.str());
                uint32_t nr = r.u32();
                for (uint32_t k = 0; k < nr && r.ok; k++) { HttpRoute hr; hr.allow = r.u32() != 0; hr.prefix = r.str(); v.routes.push_back(hr); }
                comp.http.push_back(std::move(v)); break;
            }
            default: return fail(e, CLK_EINVAL, "%s: unknown section %u", path, tag);
            }
        }
        if (!r.ok) return fail(e, CLK_EINVAL, "%s: truncated", path);
    }
    std::lock_guard<std::mutex> g(e->mu);
    if (container.size() > e->max_containers || bypass.size() > e->max_containers) return fail(e, CLK_E2BIG, "%s: %zu containers > max_entries %u", path, container.size(), e->max_containers);
    if (dns.size() > e->max_dns) return fail(e, CLK_E2BIG, "%s: %zu dns_cache rows > max_entries %u", path, dns.size(), e->max_dns);
    if (route.size() > e->max_routes) return fail(e, CLK_E2BIG, "%s: %zu routes > max_entries %u", path, route.size(), e->max_routes);
    if (cidr.size() > e->max_cidrs) return fail(e, CLK_E2BIG, "%s: %zu CIDRs > %u", path, cidr.size(), e->max_cidrs);
    if (endpoint6.size() > 256) return fail(e, CLK_E2BIG, "%s: %zu IPv6 endpoints > 256", path, endpoint6.size());
    // dns_cache is written on BOTH sides of the file: by the control-plane verbs (gc-dns, sync) into the file, and by this
    // process in memory (clk_dns_answers(apply) / clk_dns_update — dnsbpf's MapWriter).  The reference's pinned map is one
    // shared object; here the load merges: a row this process holds that the file lacks, or holds with a later expiry, stays.
    for (auto &kv : e->dns) {
        auto it = dns.find(kv.first);
        if (it == dns.end() || it->second.expire_ts < kv.second.expire_ts) { if (it == dns.end() && dns.size() >= e->max_dns) continue; dns[kv.first] = kv.second; }
    }
    e->container.swap(container); e->bypass.swap(bypass); e->dns.swap(dns); e->route.swap(route);
    e->endpoint.swap(endpoint); e->endpoint6.swap(endpoint6); e->cidr.swap(cidr); e->compiled = std::move(comp);
    e->dirty = true;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_get_stats(clk_engine *e, clk_stats *out)
try {
    if (!e || !out) return fail(e, CLK_EINVAL, "NULL");
    std::lock_guard<std::mutex> g(e->mu);
    if (!e->host_only && e->ev_recorded) {      // (the elapsed time of never-recorded events is an API error, and any error here would sit in
        float ms = 0;                          //  the runtime's last-error slot for the next launch check to find: clear what this call raises)
        if (cudaEventQuery(e->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->stats.last_kernel_ms = ms;
        else cudaGetLastError();
    }
    *out = e->stats;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_sync(clk_engine *e)
try {
    if (!e) return fail(e, CLK_EINVAL, "NULL");
    NEED_GPU(e);
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    return CLK_OK;
} CLK_GUARD_INT(e)

// ---- multi-GPU: the one collective of the path, behind the C ABI (SURVEY 8(b) clk_allgather_counters, 8(e)) -------------
// Reference analogue: metrics_map is BPF_MAP_TYPE_PERCPU_HASH (common.h:171-177) and a reader sums the per-CPU values;
// here every GPU of the host counts on its own and the sum over GPUs is an NCCL all-gather over NVLink + a local sum.
int clk_comm_unique_id(void *out, size_t cap, size_t *len)
try {
    if (!out || cap < sizeof(ncclUniqueId)) return fail(nullptr, CLK_EINVAL, "unique id needs %zu bytes", sizeof(ncclUniqueId));
    if (!nccl().h) return fail(nullptr, CLK_ENOSYS, "NCCL unavailable: %s", nccl().err.c_str());
    ncclUniqueId id;
    NCCL_TRY(nullptr, nccl().GetUniqueId(&id));
    memcpy(out, &id, sizeof id);
    if (len) *len = sizeof id;
    return CLK_OK;
} CLK_GUARD_INT(nullptr)

int clk_comm_init(clk_engine *e, const void *unique_id, size_t len, int rank, int world)
try {
    if (!e || !unique_id || len != sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world) return fail(e, CLK_EINVAL, "bad communicator arguments (id %zu bytes, rank %d of %d)", len, rank, world);
    NEED_GPU(e);
    if (!nccl().h) return fail(e, CLK_ENOSYS, "NCCL unavailable: %s", nccl().err.c_str());
    std::lock_guard<std::mutex> g(e->mu);
    if (e->comm) comm_release(e);
    ncclUniqueId id; memcpy(&id, unique_id, sizeof id);
    NCCL_TRY(e, nccl().CommInitRank(&e->comm, world, id, rank));
    e->comm_rank = rank; e->comm_world = world;
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_comm_destroy(clk_engine *e)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    CUDA_TRY(e, cudaDeviceSynchronize());
    comm_release(e);
    return CLK_OK;
} CLK_GUARD_INT(e)

int clk_comm_info(clk_engine *e, int *rank, int *world)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    std::lock_guard<std::mutex> g(e->mu);
    if (rank) *rank = e->comm_rank;
    if (world) *world = e->comm ? e->comm_world : 1;
    return CLK_OK;
} CLK_GUARD_INT(e)

// Per batch (enqueue only): dense per-container totals of this GPU's metrics_map -> all-gather of every GPU's block.
// Without a communicator (single GPU) the gathered view is this GPU's block.
int clk_allgather_counters(clk_engine *e, void *stream)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    NEED_GPU(e);
    std::lock_guard<std::mutex> g(e->mu);
    int rc = ensure_published(e); if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const uint32_t rows = 2 * e->max_containers; const size_t blk = (size_t)rows * sizeof(clk_total);
    const int world = e->comm ? e->comm_world : 1;
    if (e->T.n_crows > rows) return fail(e, CLK_E2BIG, "%u cgroup rows do not fit the %u-row totals block", e->T.n_crows, rows);
    CUDA_TRY(e, e->tot_local.ensure(blk)); CUDA_TRY(e, e->tot_all.ensure(blk * world));
    clk_total *mine = e->comm ? (clk_total *)e->tot_local.p : (clk_total *)e->tot_all.p;
    const char *err = launch_counters_totals(e->T, mine, rows, st);
    if (err) return fail(e, CLK_EIO, "totals export: %s", err);
    e->stats.kernel_launches++;
    if (e->comm) NCCL_TRY(e, nccl().AllGather(e->tot_local.p, e->tot_all.p, blk, ncclUint8, e->comm, st));
    return CLK_OK;
} CLK_GUARD_INT(e)

// The gathered totals summed by cgroup_id, rows sorted by cgroup_id (synchronises the device).
int clk_totals_read_global(clk_engine *e, clk_total *out, uint32_t cap, uint32_t *n)
try {
    if (!e || !n || (cap && !out)) return fail(e, CLK_EINVAL, "NULL");
    void *blocks; int world;
    { std::lock_guard<std::mutex> g(e->mu); blocks = e->tot_all.p; world = e->comm ? e->comm_world : 1; }
    if (!blocks) { *n = 0; return CLK_OK; }
    return clk_totals_merge(e, blocks, (uint32_t)world, out, cap, n);
} CLK_GUARD_INT(e)

// On read (enqueue only): per-key merge — compacted export of this GPU's metrics_map -> all-gather -> import into the
// merged view that clk_counters_read_global returns.
int clk_allgather_metrics(clk_engine *e, void *stream)
try {
    if (!e) return fail(e, CLK_EINVAL, "engine is NULL");
    NEED_GPU(e);
    uint32_t rows = 0; void *local, *all; int world;
    {
        std::lock_guard<std::mutex> g(e->mu);
        int rc = ensure_published(e); if (rc) return rc;
        rows = e->max_metrics + 1 + 1024; world = e->comm ? e->comm_world : 1;
        const size_t blk = (size_t)rows * sizeof(clk_metric);
        CUDA_TRY(e, e->met_local.ensure(blk)); CUDA_TRY(e, e->met_all.ensure(blk * world));
        local = e->comm ? e->met_local.p : e->met_all.p; all = e->met_all.p;
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    int rc = clk_counters_export(e, local, st); if (rc) return rc;
    if (e->comm) NCCL_TRY(e, nccl().AllGather(local, all, (size_t)rows * sizeof(clk_metric), ncclUint8, e->comm, st));
    return clk_counters_import(e, all, (uint32_t)world, st);
} CLK_GUARD_INT(e)

// Pinned host memory for packet ingest rings (true DMA for the host-buffer entry points).
int clk_host_alloc(size_t bytes, void **out)
try {
    if (!out) return fail(nullptr, CLK_EINVAL, "out is NULL");
    cudaError_t c = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (c != cudaSuccess) return fail(nullptr, CLK_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(c));
    return CLK_OK;
} CLK_GUARD_INT(nullptr)
int clk_host_free(void *p) try { if (p) cudaFreeHost(p); return CLK_OK; } CLK_GUARD_INT(nullptr)

}  // extern "C"
