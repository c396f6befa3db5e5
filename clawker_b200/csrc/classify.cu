// classify.cu — the per-packet verdict kernels (sm_100a).
//
// One thread classifies one packet: L3/L4 header parse -> socket event ->
// the reference's decision ladder (internal/ebpf/bpf/common.h:256-451, clawker.c:69-240)
// against the device snapshot -> optional name stage (CoreDNS zone / TLS SNI / DNS answers)
// -> one u32 verdict word, plus warp/CTA-aggregated metrics_map increments.
//
// Kernels
//   k_slots_fast   64-byte header slots.  2-D TMA tiles (256 slots x 64 B, SWIZZLE_64B) through an
//                  mbarrier ring into shared memory; the common case (IPv4, no options, TCP/UDP/other)
//                  runs as one branch-light predicated ladder with all 32 lanes converged; rare
//                  packets (IPv6, IP options, non-IP, malformed) are compacted into a CTA queue and
//                  classified together by the generic path.  Small tables (endpoints, dns_cache,
//                  route_map) are staged into shared memory with 1-D TMA bulk copies.
//   k_generic      any stride / varlen offsets, optional name stage; reads packets through L1.
//   k_events       pre-parsed socket events (the reference hooks' own input).
//
// This is HBM-bound integer/byte work: no tensor cores anywhere on this path.
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>
#include "dev_tables.h"
#include "classify.h"

// The warp index as a value the compiler can prove warp-uniform (a shuffle from a fixed lane): everything derived from it —
// tile / barrier / queue addresses, loop counters — can then live in uniform registers and the TMA / mbarrier sequences need no
// per-lane election loop.
#ifndef CLK_ELECT
#define CLK_ELECT 1
#endif
#ifndef CLK_UWARP
#define CLK_UWARP 1
#endif
#if CLK_UWARP
#define UNIFORM_WARP_IDX() ((uint32_t)__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0))
#else
#define UNIFORM_WARP_IDX() (threadIdx.x >> 5)
#endif
namespace clk {

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One lane of a converged warp (elect.sync): the compiler keeps the code under it on the uniform datapath (no per-lane
// election loop around UTMALDG / the mbarrier arrive) when its operands are warp-uniform.
__device__ __forceinline__ bool elect_one()
{
#if CLK_ELECT
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
    return p != 0;
#else
    return (threadIdx.x & 31u) == 0;
#endif
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void stg_stream(uint32_t *p, uint32_t v)
{
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t bswap16(uint32_t v) { return __byte_perm(v, 0, 0x4401); }   // low 16 bits swapped

// ------------------------------------------------------------------------------------------------
// per-CTA table views (shared-memory copies when staged, else the global arrays)
// ------------------------------------------------------------------------------------------------
struct Tabs {
    const uint2    *ep;    uint32_t ep_mask, ep_shift;
    const uint4    *rowq;                      // 3 x 16 B per endpoint row (dev_tables.h RowQ)
    const uint4    *dns;   uint32_t dns_shift;
    const uint4    *route; uint32_t route_shift;
    const uint32_t *pre;                       // CIDR /16 prefilter (null when there are no CIDR rules)
    const Ep6Key   *ep6;                       // IPv6 endpoints
};

#define SOCK_STREAM_ 1u
#define SOCK_DGRAM_  2u
#define SOCK_RAW_    3u
#define CLAWKER_MARK_ 0xC1A4u          // common.h:73
#define DOCKER_DNS_NBO 0x0b00007fu     // bpf_htonl(0x7f00000b), common.h:62

// dns_cache[ip] -> domain_hash.  One 16 B bucket load; the second bucket is read only when the first is
// full (slot 1 occupied) and missed.  The key 0xffffffff never enters the table (it is the empty marker):
// the host keeps that one row out of band (dns_ffff_*).
__device__ __forceinline__ bool dns_lookup(const DevTables &T, const Tabs &tb, uint32_t ip, uint32_t &hash)
{
    // One branch only (the second bucket, taken when the first is full of other keys): compares feed selects.
    bool found = false;
    if (T.dns_b) {
        uint4 v = tb.dns[hash_a(ip) >> tb.dns_shift];
        bool a = v.x == ip, b = v.z == ip;
        if (!(a | b) && v.z != 0xffffffffu) {
            v = tb.dns[hash_b(ip) >> tb.dns_shift];
            a = v.x == ip; b = v.z == ip;
        }
        found = a | b;
        if (found) hash = a ? v.y : v.w;
    }
    // the one key that looks like an empty slot is kept out of band; whatever the probe above "found" for it is void
    if (ip == 0xffffffffu) { found = T.dns_ffff_present != 0; hash = found ? T.dns_ffff_hash : 0u; }
    return found;
}

// route_map[{hash, port}] -> envoy_port.  Slot (lo, hi): hi = domain_hash, lo = dst_port << 16 | envoy_port;
// empty = ~0 (the host rejects the one route whose key and value are all ones).
__device__ __forceinline__ bool route_lookup(const DevTables &T, const Tabs &tb, uint32_t hash, uint32_t port, uint32_t &eport)
{
    if (!T.route_b) return false;
    const uint32_t f = route_fold(hash, port);
    uint4 v = tb.route[hash_a(f) >> tb.route_shift];
    bool a = v.y == hash && (v.x >> 16) == port, b = v.w == hash && (v.z >> 16) == port;
    if (!(a | b) && (v.z & v.w) != 0xffffffffu) {
        v = tb.route[hash_b(f) >> tb.route_shift];
        a = v.y == hash && (v.x >> 16) == port; b = v.w == hash && (v.z >> 16) == port;
    }
    const uint32_t lo = a ? v.x : (b ? v.z : 0xffffffffu);
    if (lo == 0xffffffffu) return false;           // miss, or an empty slot "matching" {0xffffffff, 65535}
    eport = lo & 0xffffu;
    return true;
}

// A12: dst inside any allowed CIDR.  2-bit /24 directory + 256-bit leaf bitmaps, both L2-resident.
__device__ __forceinline__ bool cidr_hit(const DevTables &T, const Tabs &tb, uint32_t ip)
{
    if (T.n_cidr == 0) return false;
    uint32_t be = __byte_perm(ip, 0, 0x0123);          // host-order address
    if (!((tb.pre[be >> 21] >> ((be >> 16) & 31u)) & 1u)) return false;       // no rule touches this /16
    uint32_t blk = be >> 8;
    uint32_t two = (__ldg(&T.cidr_l1[blk >> 4]) >> ((blk & 15u) * 2u)) & 3u;
    if (two & 1u) return true;
    if (two & 2u) {
        uint32_t h = hash_a(blk) >> T.cidr_l2_shift;
        #pragma unroll 1
        for (;;) {
            uint32_t k = __ldg(&T.cidr_l2_keys[h]);
            if (k == 0) return false;
            if (k == blk + 1) return (__ldg(&T.cidr_l2_bits[h * 8 + ((be & 255u) >> 5)]) >> (be & 31u)) & 1u;
            h = (h + 1) & T.cidr_l2_mask;
        }
    }
    return false;
}

// endpoint by IPv4 address -> EpRow index, or -1
__device__ __forceinline__ int ep_find(const DevTables &T, const Tabs &tb, uint32_t ip)
{
    if (!T.ep_slots) return -1;
    uint32_t h = hash_a(ip) >> tb.ep_shift;
    #pragma unroll 1
    for (;;) {
        uint2 s = tb.ep[h];
        if (s.y == 0) return -1;
        if (s.x == ip) return (int)s.y - 1;
        h = (h + 1) & tb.ep_mask;
    }
}
__device__ __forceinline__ int ep_find6(const DevTables &T, const Tabs &tb, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3)
{
    if (!T.ep6) return -1;
    uint32_t h = hash_a(a0 ^ a1 ^ a2 ^ a3) >> T.ep6_shift;
    #pragma unroll 1
    for (;;) {
        const uint4 k = *(const uint4 *)&tb.ep6[h];
        const uint32_t row = tb.ep6[h].row;
        if (row == 0xffffffffu) return -1;
        if (k.x == a0 && k.y == a1 && k.z == a2 && k.w == a3) return (int)row;
        h = (h + 1) & T.ep6_mask;
    }
}

// ------------------------------------------------------------------------------------------------
// decision ladder (generic form: used by the rare-packet path, the varlen kernel and the events kernel)
// ------------------------------------------------------------------------------------------------
struct Decision {
    uint32_t verdict, metric;     // CLK_V_*, CLK_M_*
    uint32_t new_ip, new_port;    // valid when verdict == REWRITE
    uint32_t m_hash, m_port;      // metric key parts (valid when metric != 0)
};

__device__ __forceinline__ void emit(Decision &d, uint32_t hash, uint32_t port, uint32_t action)
{
    d.metric = action + 1; d.m_hash = hash; d.m_port = port;
}

// decide_connect, common.h:323-410 (+ A12 at :350)
__device__ __forceinline__ void decide_connect(const DevTables &T, const Tabs &tb, const clk_container_config *cfg,
                                               uint32_t mark, uint32_t type, uint32_t ip, uint32_t port, Decision &d)
{
    if (mark == CLAWKER_MARK_) return;
    if (port == 53) { d.verdict = CLK_V_REWRITE; d.new_ip = cfg->coredns_ip; d.new_port = 53; emit(d, 0, 53, CLK_ACTION_ALLOW); return; }
    if ((ip & 0xffu) == 0x7fu) return;                                        // is_loopback :213-216
    if ((ip & cfg->net_mask) == cfg->net_addr) return;                        // is_in_subnet :220-223
    if (cidr_hit(T, tb, ip)) return;
    uint32_t pp = cfg->host_proxy_port;
    if (cfg->host_proxy_ip != 0 && ip == cfg->host_proxy_ip && port == pp) return;
    if (ip == cfg->gateway_ip) {
        if (pp != 0 && port == pp) return;
        d.verdict = CLK_V_REWRITE; d.new_ip = cfg->envoy_ip; d.new_port = cfg->egress_port;
        emit(d, 0, port, CLK_ACTION_DENY); return;
    }
    if (type == SOCK_DGRAM_) { emit(d, 0, port, CLK_ACTION_DENY); d.verdict = CLK_V_DENY; return; }
    uint32_t hash = 0, eport;
    if (dns_lookup(T, tb, ip, hash)) {
        if (route_lookup(T, tb, hash, port, eport)) {
            d.verdict = CLK_V_REWRITE; d.new_ip = cfg->envoy_ip; d.new_port = eport;
            emit(d, hash, port, CLK_ACTION_ALLOW); return;
        }
    }
    d.verdict = CLK_V_REWRITE; d.new_ip = cfg->envoy_ip; d.new_port = cfg->egress_port;
    emit(d, hash, port, CLK_ACTION_ALLOW);
}

// decide_sendmsg, common.h:416-439 (+ A12)
__device__ __forceinline__ void decide_sendmsg(const DevTables &T, const Tabs &tb, const clk_container_config *cfg, uint32_t ip, uint32_t port, Decision &d)
{
    if (port == 53) { d.verdict = CLK_V_REWRITE; d.new_ip = cfg->coredns_ip; d.new_port = 53; return; }
    if ((ip & 0xffu) == 0x7fu) return;
    if ((ip & cfg->net_mask) == cfg->net_addr) return;
    if (cidr_hit(T, tb, ip)) return;
    emit(d, 0, port, CLK_ACTION_DENY); d.verdict = CLK_V_DENY;
}

// The seven programs (clawker.c:69-240) behind enter_enforced (common.h:256-280).
// ip = user_ip4 for *4 hooks, user_ip6[3] for *6 hooks; a0..a2 = user_ip6[0..2].
__device__ __forceinline__ void run_hooks(const DevTables &T, const Tabs &tb, uint32_t hook, uint32_t type, uint32_t uid,
                                          uint32_t mark, uint32_t flags, const clk_container_config *cfg,
                                          uint32_t ip, uint32_t port, uint32_t a0, uint32_t a1, uint32_t a2, Decision &d)
{
    if (hook == CLK_HOOK_NONE) return;
    if (hook == CLK_HOOK_CONNECT4 && type != SOCK_STREAM_ && type != SOCK_DGRAM_) return;     // clawker.c:72
    if (uid == 0) return;                                                                      // common.h:259-261
    bool recv = hook == CLK_HOOK_RECVMSG4 || hook == CLK_HOOK_RECVMSG6;
    if (!recv && (flags & CLK_ROW_BYPASS)) { emit(d, 0, 0, CLK_ACTION_BYPASS); return; }       // :265-271
    if (!(flags & CLK_ROW_ENFORCED)) return;                                                   // :273-275
    switch (hook) {
    case CLK_HOOK_CONNECT4: decide_connect(T, tb, cfg, mark, type, ip, port, d); return;
    case CLK_HOOK_SENDMSG4: decide_sendmsg(T, tb, cfg, ip, port, d); return;
    case CLK_HOOK_RECVMSG4:
        if (ip == cfg->coredns_ip && port == 53) { d.verdict = CLK_V_REWRITE; d.new_ip = DOCKER_DNS_NBO; d.new_port = 53; }
        return;
    case CLK_HOOK_CONNECT6:
    case CLK_HOOK_SENDMSG6:
        if (a0 == 0 && a1 == 0 && a2 == 0 && ip == 0x01000000u) return;                        // ::1
        if (a0 == 0 && a1 == 0 && a2 == 0xffff0000u) {                                         // ::ffff:a.b.c.d
            if (hook == CLK_HOOK_CONNECT6) {
                if (type != SOCK_STREAM_ && type != SOCK_DGRAM_) return;                       // clawker.c:146
                decide_connect(T, tb, cfg, mark, type, ip, port, d);
            } else decide_sendmsg(T, tb, cfg, ip, port, d);
            return;
        }
        emit(d, 0, port, CLK_ACTION_DENY); d.verdict = CLK_V_DENY;                             // clawker.c:155,186
        return;
    case CLK_HOOK_RECVMSG6:
        if (!(a0 == 0 && a1 == 0 && a2 == 0xffff0000u)) return;
        if (ip == cfg->coredns_ip && port == 53) { d.verdict = CLK_V_REWRITE; d.new_ip = DOCKER_DNS_NBO; d.new_port = 53; }
        return;
    case CLK_HOOK_SOCK_CREATE:
        if (type == SOCK_RAW_) { emit(d, 0, 0, CLK_ACTION_DENY); d.verdict = CLK_V_DENY; }      // clawker.c:235-238
        return;
    }
}

__device__ __forceinline__ uint32_t target_of(uint32_t coredns_ip, uint32_t envoy_ip, uint32_t egress_port, uint32_t ip, uint32_t port)
{
    if (ip == coredns_ip && port == 53) return CLK_T_COREDNS;
    if (ip == envoy_ip && port == egress_port) return CLK_T_EGRESS;
    if (ip == envoy_ip) return CLK_T_TCP;
    return CLK_T_NONE;
}

__device__ __forceinline__ uint32_t make_word(const Decision &d, const clk_container_config *cfg, bool has_cfg,
                                              uint32_t hook, uint32_t name, uint32_t nupd)
{
    uint32_t w = d.verdict | (d.metric << 2) | (name << 4) | (hook << 24) | (nupd << 27);
    if (d.verdict == CLK_V_REWRITE) {
        if (has_cfg) w |= target_of(cfg->coredns_ip, cfg->envoy_ip, cfg->egress_port, d.new_ip, d.new_port) << 6;
        w |= (d.new_port & 0xffffu) << 8;
    }
    return w;
}

// ------------------------------------------------------------------------------------------------
// metrics: warp match -> CTA shared-memory table (hot keys) -> device table (metric_inc common.h:190-210)
// ------------------------------------------------------------------------------------------------
#define MSLOTS 512u              // default CTA aggregation table; k_slots_fast sizes its own from the shared memory left

__device__ __forceinline__ uint32_t mhash(unsigned long long key)
{
    uint32_t t = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA77u);
    return t ^ (t >> 15);
}

// Probe / insert / drop: everything except "the key sits in its home slot" (kept out of line so the hot
// path stays two instructions long).  Returns the increments it had to drop.
__device__ __noinline__ unsigned int metric_global_slow(const DevTables &T, unsigned long long key, unsigned int cnt, uint32_t h)
{
    // A new key is a BPF_NOEXIST insert, refused once the map holds max_entries (metric_inc common.h:206-209).  The
    // bound is strict, as in the kernel's hash map: an insert first takes a ticket from the key counter and gives it
    // back if it loses the slot to the same key or finds the map full.  The merged view (import) is uncapped.
    const bool capped = T.m_max != 0xffffffffu;
    bool ticket = false; unsigned int lost = cnt;
    #pragma unroll 1
    for (int p = 0; p < 64; p++) {
        unsigned long long cur = T.m_keys[h];
        if (cur != key) {
            if (cur != 0ull) { h = (h + 1) & T.m_mask; continue; }
            if (capped && !ticket) {
                if (*(volatile unsigned long long *)&T.m_dropped[1] >= T.m_max) break;
                if (atomicAdd(&T.m_dropped[1], 1ull) >= T.m_max) { atomicAdd(&T.m_dropped[1], ~0ull); break; }
                ticket = true;
            }
            cur = atomicCAS(&T.m_keys[h], 0ull, key);
            if (cur == 0ull) ticket = false;                            // inserted: the ticket is spent
            else if (cur != key) { h = (h + 1) & T.m_mask; continue; }
        }
        atomicAdd(&T.m_counts[h], (unsigned long long)cnt);
        lost = 0u;
        break;
    }
    if (ticket) atomicAdd(&T.m_dropped[1], ~0ull);
    return lost;
}

__device__ __forceinline__ unsigned int metric_global_add(const DevTables &T, unsigned long long key, unsigned int cnt)
{
    const uint32_t h = mhash(key) & T.m_mask;
    if (T.m_keys[h] == key) { atomicAdd(&T.m_counts[h], (unsigned long long)cnt); return 0u; }
    return metric_global_slow(T, key, cnt, h);
}

// Per-CTA aggregation state.  Increments are not applied by the (few, divergent) lanes that produce them:
// each warp appends its distinct (key, count) pairs to a small log in shared memory and applies the log 32
// entries at a time with every lane busy — first against the CTA's table of frequent keys, and only the misses
// against the device table (so L2 atomics are issued 32 at a time instead of 1-3 at a time).
#define MLOG_CAP 64u
struct MetricAgg {
    unsigned long long *keys; unsigned int *counts; uint32_t mask;     // CTA table (shared memory)
    unsigned long long *log_keys; unsigned int *log_cnts; uint32_t log_n;   // this warp's log
    unsigned int dropped;
};

// Two probes at most: keys that find a slot early (the frequent ones, by construction) stay in shared
// memory for the CTA's lifetime; everything else goes to the device table.
__device__ __forceinline__ bool metric_cta_try(MetricAgg &ag, unsigned long long key, unsigned int cnt)
{
    uint32_t h = mhash(key) & ag.mask;
    #pragma unroll
    for (int p = 0; p < 2; p++) {
        unsigned long long cur = ag.keys[h];
        if (cur == 0ull) cur = atomicCAS(&ag.keys[h], 0ull, key);
        if (cur == 0ull || cur == key) { atomicAdd(&ag.counts[h], cnt); return true; }
        h = (h + 1) & ag.mask;
    }
    return false;
}

// Apply the first `take` (<= 32) log entries to the device table, one per lane.  All 32 lanes must call.
__device__ __forceinline__ void metric_log_apply(const DevTables &T, MetricAgg &ag, uint32_t take)
{
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long k = 0ull, k2 = 0ull; unsigned int c = 0u, c2 = 0u;
    if (lane < take) { k = ag.log_keys[lane]; c = ag.log_cnts[lane]; }
    if (lane + 32 < ag.log_n) { k2 = ag.log_keys[lane + 32]; c2 = ag.log_cnts[lane + 32]; }
    __syncwarp();
    if (lane + 32 < ag.log_n) { ag.log_keys[lane] = k2; ag.log_cnts[lane] = c2; }
    ag.log_n -= take;
    if (lane < take) ag.dropped += metric_global_add(T, k, c);
    __syncwarp();
}

// All 32 lanes must call (tail lanes with has == false).  Every lane adds its own increment to the CTA table
// (shared-memory atomics serialise equal addresses in the memory pipe, not in the instruction stream); only the
// increments that find no slot there are logged, and the log goes to the device table 32 entries at a time.
__device__ __forceinline__ void metric_warp_emit(const DevTables &T, MetricAgg &ag, bool has, unsigned long long key)
{
    const uint32_t lane = threadIdx.x & 31u;
#ifdef CLK_METRIC_MATCH
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (m == 0) return;
    unsigned peers = 0;
    if (has) peers = __match_any_sync(m, key);
    const bool leader = has && (int)lane == __ffs(peers) - 1;
    const unsigned int cnt = __popc(peers);
    const bool miss = leader && !metric_cta_try(ag, key, cnt);
#else
    const unsigned int cnt = 1u;
    const bool miss = has && !metric_cta_try(ag, key, 1u);
#endif
    const unsigned mm = __ballot_sync(0xffffffffu, miss);
    if (mm == 0) return;
    if (miss) { const uint32_t pos = ag.log_n + __popc(mm & ((1u << lane) - 1u)); ag.log_keys[pos] = key; ag.log_cnts[pos] = cnt; }
    ag.log_n += __popc(mm);
    __syncwarp();
    if (ag.log_n >= 32) metric_log_apply(T, ag, 32);
}

__device__ __forceinline__ void metric_cta_flush(const DevTables &T, MetricAgg &ag)
{
    if (ag.log_n) metric_log_apply(T, ag, ag.log_n);                    // log_n < 32 here
    __syncthreads();
    for (uint32_t i = threadIdx.x; i <= ag.mask; i += blockDim.x)
        if (ag.keys[i]) ag.dropped += metric_global_add(T, ag.keys[i], ag.counts[i]);
    unsigned long long d = ag.dropped;                                  // one atomic per warp, and only if something was dropped
    for (int o = 16; o; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31u) == 0 && d) atomicAdd(&T.m_dropped[0], d);
}

// ------------------------------------------------------------------------------------------------
// generic header parse: 64 bytes in four 16 B registers -> event
// ------------------------------------------------------------------------------------------------
struct Parsed {
    uint32_t hook, type, ip, port, a0, a1, a2;
    uint32_t proto, l4off, end;        // for the name stage (varlen kernel)
    bool v6, ingress, ans_candidate;   // ans_candidate: UDP sport 53 to a CoreDNS address, no endpoint
    const EpRow *row;
};

__device__ __forceinline__ uint32_t hdr_byte(const uint4 &q0, const uint4 &q1, const uint4 &q2, const uint4 &q3, uint32_t i)
{
    // rare path (IPv4 options): dynamic byte index into the 64 header bytes
    uint32_t w[16] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w };
    return (w[(i >> 2) & 15] >> ((i & 3) * 8)) & 0xffu;
}

__device__ __forceinline__ bool is_coredns(const DevTables &T, uint32_t ip)
{
    for (uint32_t i = 0; i < T.n_coredns; i++) if (T.coredns_ips[i] == ip) return true;
    return false;
}

// Second half of the IPv4 parse once addresses and ports are known.
__device__ __forceinline__ void finish_v4(const DevTables &T, const Tabs &tb, uint32_t flags, Parsed &p,
                                          uint32_t src, uint32_t dst, uint32_t proto, uint32_t sport, uint32_t dport)
{
    int r = ep_find(T, tb, src);
    bool ingress = false;
    if (r < 0 && proto == 17) { r = ep_find(T, tb, dst); ingress = r >= 0; }
    if (r < 0) {
        if (proto == 17 && sport == 53 && (flags & CLK_F_NAME_STAGE) && is_coredns(T, dst)) p.ans_candidate = true;
        return;
    }
    p.row = &T.ep_rows[r]; p.ingress = ingress; p.ip = ingress ? src : dst;
    if (proto == 6)       { p.hook = CLK_HOOK_CONNECT4; p.type = SOCK_STREAM_; p.port = dport; }
    else if (proto == 17) { p.type = SOCK_DGRAM_; if (ingress) { p.hook = CLK_HOOK_RECVMSG4; p.port = sport; } else { p.hook = CLK_HOOK_SENDMSG4; p.port = dport; } }
    else                  { p.hook = CLK_HOOK_SOCK_CREATE; p.type = SOCK_RAW_; }
}

// pk: the packet in global memory (needed only when IPv4 options push the L4 ports past byte 64).
// avail = captured bytes of this packet.
__device__ __forceinline__ void parse_headers(const DevTables &T, const Tabs &tb, const uint4 &q0, const uint4 &q1, const uint4 &q2,
                                              const uint4 &q3, const uint8_t *pk, uint32_t avail, uint32_t flags, Parsed &p)
{
    p.hook = CLK_HOOK_NONE; p.type = 0; p.ip = 0; p.port = 0; p.a0 = p.a1 = p.a2 = 0;
    p.proto = 0; p.l4off = 0; p.end = 0; p.v6 = false; p.ingress = false; p.ans_candidate = false; p.row = nullptr;
    if (avail < 14) return;
    uint32_t et = bswap16(q0.w);
    uint32_t sport = 0, dport = 0;
    if (et == 0x0800u) {
        uint32_t vi = (q0.w >> 16) & 0xffu;
        if (avail < 34 || (vi >> 4) != 4) return;
        uint32_t ihl = vi & 15u;
        if (ihl < 5) return;
        if (bswap16(q1.y) & 0x1fffu) return;                       // non-first fragment
        uint32_t l4off = 14 + 4 * ihl;
        uint32_t proto = q1.y >> 24;
        uint32_t end = 14 + bswap16(q1.x); if (end > avail) end = avail;
        uint32_t src = __byte_perm(q1.z, q1.w, 0x5432), dst = __byte_perm(q1.w, q2.x, 0x5432);
        if (proto == 6 || proto == 17) {
            if (l4off + 4 > avail) return;
            if (ihl == 5) { sport = bswap16(q2.x >> 16); dport = bswap16(q2.y); }
            else if (l4off + 4 <= 64) {
                sport = (hdr_byte(q0, q1, q2, q3, l4off) << 8) | hdr_byte(q0, q1, q2, q3, l4off + 1);
                dport = (hdr_byte(q0, q1, q2, q3, l4off + 2) << 8) | hdr_byte(q0, q1, q2, q3, l4off + 3);
            } else {
                sport = ((uint32_t)__ldg(pk + l4off) << 8) | __ldg(pk + l4off + 1);
                dport = ((uint32_t)__ldg(pk + l4off + 2) << 8) | __ldg(pk + l4off + 3);
            }
        }
        p.proto = proto; p.l4off = l4off; p.end = end;
        finish_v4(T, tb, flags, p, src, dst, proto, sport, dport);
    } else if (et == 0x86DDu) {
        uint32_t vi = (q0.w >> 16) & 0xffu;
        if (avail < 54 || (vi >> 4) != 6) return;
        uint32_t proto = q1.y & 0xffu;
        uint32_t end = 54 + bswap16(q1.x >> 16); if (end > avail) end = avail;
        if (proto == 6 || proto == 17) {
            if (58 > avail) return;
            sport = bswap16(q3.y >> 16); dport = bswap16(q3.z);
        }
        uint32_t s0 = __byte_perm(q1.y, q1.z, 0x5432), s1 = __byte_perm(q1.z, q1.w, 0x5432),
                 s2 = __byte_perm(q1.w, q2.x, 0x5432), s3 = __byte_perm(q2.x, q2.y, 0x5432);
        uint32_t d0 = __byte_perm(q2.y, q2.z, 0x5432), d1 = __byte_perm(q2.z, q2.w, 0x5432),
                 d2 = __byte_perm(q2.w, q3.x, 0x5432), d3 = __byte_perm(q3.x, q3.y, 0x5432);
        int r = ep_find6(T, tb, s0, s1, s2, s3);
        bool ingress = false;
        if (r < 0 && proto == 17) { r = ep_find6(T, tb, d0, d1, d2, d3); ingress = r >= 0; }
        p.proto = proto; p.l4off = 54; p.end = end; p.v6 = true;
        if (r < 0) return;
        p.row = &T.ep_rows[r]; p.ingress = ingress;
        if (ingress) { p.a0 = s0; p.a1 = s1; p.a2 = s2; p.ip = s3; } else { p.a0 = d0; p.a1 = d1; p.a2 = d2; p.ip = d3; }
        if (proto == 6)       { p.hook = CLK_HOOK_CONNECT6; p.type = SOCK_STREAM_; p.port = dport; }
        else if (proto == 17) { p.type = SOCK_DGRAM_; if (ingress) { p.hook = CLK_HOOK_RECVMSG6; p.port = sport; } else { p.hook = CLK_HOOK_SENDMSG6; p.port = dport; } }
        else                  { p.hook = CLK_HOOK_SOCK_CREATE; p.type = SOCK_RAW_; }
    }
}

#include "name_stage.cuh"

// Classify one parsed packet (no name stage) -> decision + metric key.  Returns 0 when there is no event.
__device__ __forceinline__ uint32_t classify_parsed(const DevTables &T, const Tabs &tb, const Parsed &p, Decision &d,
                                                    bool &has_metric, unsigned long long &mkey)
{
    d.verdict = 0; d.metric = 0; d.new_ip = 0; d.new_port = 0; d.m_hash = 0; d.m_port = 0;
    has_metric = false; mkey = 0;
    if (p.hook == CLK_HOOK_NONE) return 0;
    const EpRow *r = p.row;
    run_hooks(T, tb, p.hook, p.type, r->uid, r->so_mark, r->flags, &r->cfg, p.ip, p.port, p.a0, p.a1, p.a2, d);
    if (d.verdict == CLK_V_REWRITE && d.new_ip == p.ip && d.new_port == p.port) d.verdict = CLK_V_PASS;   // identity rewrite
    if (d.metric) { has_metric = true; mkey = metric_pack(d.m_hash, d.m_port, r->crow, d.metric - 1); }
    return 1;
}

// One packet through the generic path (any layout; optional name stage).  kName = false: the instantiation the
// header-only kernels call (no payload parser is compiled into them: code size, registers and stack of their hot loops).
template <bool kName>
__device__ __noinline__ uint32_t generic_one(const DevTables &T, const Tabs &tb, const uint8_t *pk, uint32_t avail, uint32_t flags,
                                             bool &hm, unsigned long long &mk)
{
    uint4 z = make_uint4(0, 0, 0, 0), q0 = z, q1 = z, q2 = z, q3 = z;
    const uint4 *pq = (const uint4 *)pk;
    if (avail >= 16) q0 = __ldg(pq);
    if (avail >= 32) q1 = __ldg(pq + 1);
    if (avail >= 48) q2 = __ldg(pq + 2);
    if (avail >= 64) q3 = __ldg(pq + 3);
    Parsed p; Decision d;
    parse_headers(T, tb, q0, q1, q2, q3, pk, avail, flags, p);
    if (kName && p.ans_candidate) { hm = false; mk = 0; return dns_answer_word(T, pk, p); }
    if (!classify_parsed(T, tb, p, d, hm, mk)) return 0;
    uint32_t name = CLK_N_NA;
    if (kName && (flags & CLK_F_NAME_STAGE)) name = name_stage(T, pk, avail, p, d);
    return make_word(d, &p.row->cfg, (p.row->flags & CLK_ROW_ENFORCED) != 0, p.hook, name, 0);
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up common to the kernels
// ------------------------------------------------------------------------------------------------
struct SmemLayout { uint32_t tiles, agg_keys, agg_counts, mlog, ep, soa, dns, route, pre, ep6, queue, bars, total; };

__host__ __device__ inline SmemLayout smem_layout(const DevTables &T, uint32_t tile_bytes, uint32_t queue_bytes, uint32_t n_bars, uint32_t agg_slots = MSLOTS, uint32_t n_warps = 8)
{
    SmemLayout L; uint32_t o = 0;
    L.tiles = o; o += tile_bytes;                          // 1024 B aligned tiles first
    L.agg_keys = o; o += agg_slots * 8;
    L.agg_counts = o; o += agg_slots * 4;
    L.mlog = o; o += n_warps * MLOG_CAP * 12;
    L.ep = o; o += (T.stage_ep_bytes + 15u) & ~15u;
    L.soa = o; o += (T.stage_soa_bytes + 15u) & ~15u;
    L.dns = o; o += (T.stage_dns_bytes + 15u) & ~15u;
    L.route = o; o += (T.stage_route_bytes + 15u) & ~15u;
    L.pre = o; o += (T.stage_pre_bytes + 15u) & ~15u;
    L.ep6 = o; o += (T.stage_ep6_bytes + 15u) & ~15u;
    L.queue = o; o += queue_bytes;
    L.bars = o; o += (n_bars + 1) * 8 + 16;                // + a CTA-wide counter word after the barriers
    L.total = o;
    return L;
}

// Stage the small tables into shared memory with 1-D TMA bulk copies (thread 0), zero the metric
// aggregation table, and return the table views.  bars[n_bars] is used for the table copy.
// kAllStaged: the host guarantees every small table is staged, so the views are shared-memory pointers
// unconditionally (the compiler can then address them as shared: LDS with 32-bit addresses, fewer registers).
template <bool kAllStaged>
__device__ __forceinline__ void cta_prologue(const DevTables &T, uint8_t *smem, const SmemLayout &L, uint32_t n_bars, Tabs &tb, MetricAgg &ag, uint32_t agg_slots = MSLOTS)
{
    ag.keys = (unsigned long long *)(smem + L.agg_keys);
    ag.counts = (unsigned int *)(smem + L.agg_counts); ag.dropped = 0u; ag.mask = agg_slots - 1;
    const uint32_t wu = UNIFORM_WARP_IDX();
    ag.log_keys = (unsigned long long *)(smem + L.mlog) + wu * MLOG_CAP;
    ag.log_cnts = (unsigned int *)(smem + L.mlog + (blockDim.x >> 5) * MLOG_CAP * 8) + wu * MLOG_CAP; ag.log_n = 0;
    for (uint32_t i = threadIdx.x; i < agg_slots; i += blockDim.x) { ag.keys[i] = 0ull; ag.counts[i] = 0u; }
    uint64_t *bars = (uint64_t *)(smem + L.bars);
    tb.ep = T.ep_slots; tb.ep_mask = T.ep_mask; tb.ep_shift = T.ep_shift; tb.rowq = T.ep_rowq;
    tb.dns = T.dns_b; tb.dns_shift = T.dns_shift; tb.route = T.route_b; tb.route_shift = T.route_shift; tb.pre = T.cidr_pre; tb.ep6 = T.ep6;
    const uint32_t nb = T.stage_ep_bytes + T.stage_soa_bytes + T.stage_dns_bytes + T.stage_route_bytes + T.stage_pre_bytes + T.stage_ep6_bytes;
    if (threadIdx.x == 0) { for (uint32_t i = 0; i <= n_bars; i++) mbar_init(&bars[i], 1); *(uint32_t *)&bars[n_bars + 1] = 0; fence_mbar_init(); }
    __syncthreads();
    if (nb) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bars[n_bars], nb);
            if (T.stage_ep_bytes)    tma_load_1d(smem + L.ep, T.ep_slots, T.stage_ep_bytes, &bars[n_bars]);
            if (T.stage_soa_bytes)   tma_load_1d(smem + L.soa, T.ep_rowq, T.stage_soa_bytes, &bars[n_bars]);
            if (T.stage_dns_bytes)   tma_load_1d(smem + L.dns, T.dns_b, T.stage_dns_bytes, &bars[n_bars]);
            if (T.stage_route_bytes) tma_load_1d(smem + L.route, T.route_b, T.stage_route_bytes, &bars[n_bars]);
            if (T.stage_pre_bytes)   tma_load_1d(smem + L.pre, T.cidr_pre, T.stage_pre_bytes, &bars[n_bars]);
            if (T.stage_ep6_bytes)   tma_load_1d(smem + L.ep6, T.ep6, T.stage_ep6_bytes, &bars[n_bars]);
        }
        mbar_wait(&bars[n_bars], 0);
        if (kAllStaged) {
            tb.ep = (const uint2 *)(smem + L.ep); tb.rowq = (const uint4 *)(smem + L.soa); tb.dns = (const uint4 *)(smem + L.dns);
            tb.route = (const uint4 *)(smem + L.route); tb.pre = (const uint32_t *)(smem + L.pre); tb.ep6 = (const Ep6Key *)(smem + L.ep6);
            return;
        }
        if (T.stage_ep_bytes)    tb.ep = (const uint2 *)(smem + L.ep);
        if (T.stage_soa_bytes)   tb.rowq = (const uint4 *)(smem + L.soa);
        if (T.stage_dns_bytes)   tb.dns = (const uint4 *)(smem + L.dns);
        if (T.stage_route_bytes) tb.route = (const uint4 *)(smem + L.route);
        if (T.stage_pre_bytes)   tb.pre = (const uint32_t *)(smem + L.pre);
        if (T.stage_ep6_bytes)   tb.ep6 = (const Ep6Key *)(smem + L.ep6);
    }
}

// ------------------------------------------------------------------------------------------------
// kernel 1: 64-byte header slots through a TMA tile ring  (BASELINE configs 2, 4, 5)
// ------------------------------------------------------------------------------------------------
#define FAST_WARPS 32           // warps per CTA: one CTA per SM, so tables are staged once and the metric cache is large
#define WTILE_BYTES 2048        // one warp tile: 32 slots x 64 B
#define WQ_CAP 64               // per-warp queue of rare packets (drained 32 at a time)

// The common case as one predicated ladder.  Inputs: bytes 12..15 (w3), 16..31 (q1), 32..39 (q2x, q2y).
// Returns false when the packet is not "simple IPv4" and must take the generic path.
// nm (name-stage work this packet still needs): 0 none, NM_DNSQ / NM_DNSQ_TCP / NM_TLS / NM_DNSA.
#define NM_NONE 0u
#define NM_DNSQ 1u
#define NM_DNSQ_TCP 2u
#define NM_TLS 3u
#define NM_DNSA 4u
#define NM_HTTP 5u         // second-level queue: filled when the egress-payload queue is drained (first payload byte != 0x16)
// The decision on the five header fields alone (what a compact ingest record carries, clk_hdr16): endpoint lookup,
// enter_enforced + decide_connect / decide_sendmsg / recvmsg / sock_create as one predicated ladder.
// end / doff (kName only): end of the IP datagram inside the capture and the TCP data offset, for the payload window.
template <bool kName>
__device__ __forceinline__ void fast_fields(const DevTables &T, const Tabs &tb, uint32_t proto, uint32_t src, uint32_t dst, uint32_t sport, uint32_t dport,
                                            uint32_t end, uint32_t doff, uint32_t flags, uint32_t &word, bool &hm, unsigned long long &mk, uint32_t &nm)
{
    const bool tcp = proto == 6, udp = proto == 17, raw = !tcp && !udp;

    int r = ep_find(T, tb, src);
    bool ingress = false;
    if (r < 0 && udp) { r = ep_find(T, tb, dst); ingress = r >= 0; }
    word = 0; hm = false; mk = 0; nm = NM_NONE;
    if (r < 0) {                                                                 // no endpoint: no socket event ...
        if (kName && (flags & CLK_F_NAME_STAGE) && udp && sport == 53 && is_coredns(T, dst)) nm = NM_DNSA;   // ... but maybe an upstream answer reaching CoreDNS
        return;
    }

    const uint4 *rq = tb.rowq + 3 * r;
    const uint4 ra = rq[0], rb = rq[1], rc = rq[2];
    const uint32_t uid = ra.x, mark = ra.y, fl = ra.z, crow = ra.w;
    const uint32_t envoy = rb.x, coredns = rb.y, gateway = rb.z, hp_ip = rb.w;
    const uint32_t net = rc.x, mask = rc.y, hp_port = rc.z & 0xffffu, egress_port = rc.z >> 16;

    const uint32_t ip = ingress ? src : dst, port = raw ? 0u : (ingress ? sport : dport);
    const uint32_t hook = tcp ? CLK_HOOK_CONNECT4 : udp ? (ingress ? CLK_HOOK_RECVMSG4 : CLK_HOOK_SENDMSG4) : CLK_HOOK_SOCK_CREATE;
    const bool recv = udp && ingress;
    const bool act = uid != 0;                                                    // common.h:259-261
    const bool byp = (fl & CLK_ROW_BYPASS) != 0, enf = (fl & CLK_ROW_ENFORCED) != 0;
    const bool m_byp = act && !recv && byp;                                       // :265-271
    const bool live = act && enf && !m_byp;                                       // :273-279

    uint32_t verdict = CLK_V_PASS, metric = m_byp ? CLK_M_BYPASS : CLK_M_NONE, nip = 0, nport = 0, mh = 0, mp = 0;
    // The same ladder with its cheap outcomes as predicated assignments; only the CIDR directory walk and the dns_cache /
    // route_map probes are branches.
    const bool go = live && !raw && !recv && !(tcp && mark == CLAWKER_MARK_);            // decide_connect / decide_sendmsg body (:332-335: connect only)
    const bool is53 = port == 53;
    if (live && raw) { verdict = CLK_V_DENY; metric = CLK_M_DENY; }                       // sock_create, clawker.c:235-238
    if (live && recv && ip == coredns && is53) { verdict = CLK_V_REWRITE; nip = DOCKER_DNS_NBO; nport = 53; }   // recvmsg4, clawker.c:119-122
    if (go && is53) { verdict = CLK_V_REWRITE; nip = coredns; nport = 53; if (tcp) { metric = CLK_M_ALLOW; mp = 53; } }   // :339-345 / :423-428
    bool ext = go && !is53 && (ip & 0xffu) != 0x7fu && (ip & mask) != net;               // :347-351
    if (ext) ext = !cidr_hit(T, tb, ip);                                                  // (+A12)
    const bool hp = ext && tcp && hp_ip != 0 && ip == hp_ip && port == hp_port;           // :353-355
    const bool gw = ext && tcp && !hp && ip == gateway;                                   // :362-370
    if (ext && udp) { verdict = CLK_V_DENY; metric = CLK_M_DENY; mp = port; }             // decide_sendmsg :436-437
    if (gw && !(hp_port != 0 && port == hp_port)) { verdict = CLK_V_REWRITE; nip = envoy; nport = egress_port; metric = CLK_M_DENY; mp = port; }
    if (ext && tcp && !hp && !gw) {                                                       // :386-409
        uint32_t hash = 0, eport = egress_port;
        if (dns_lookup(T, tb, ip, hash)) route_lookup(T, tb, hash, port, eport);
        verdict = CLK_V_REWRITE; nip = envoy; nport = eport; metric = CLK_M_ALLOW; mh = hash; mp = port;
    }
    if (verdict == CLK_V_REWRITE && nip == ip && nport == port) verdict = CLK_V_PASS;          // identity rewrite
    uint32_t w = verdict | (metric << 2) | (hook << 24);
    if (verdict == CLK_V_REWRITE) {
        if (enf) w |= target_of(coredns, envoy, egress_port, nip, nport) << 6;
        w |= (nport & 0xffffu) << 8;
    }
    word = w;
    if (kName && (flags & CLK_F_NAME_STAGE) && enf && !ingress && !raw && verdict != CLK_V_DENY) {
        const uint32_t fip = verdict == CLK_V_REWRITE ? nip : ip, fport = verdict == CLK_V_REWRITE ? nport : port;
        const uint32_t po = 34 + doff;
        const bool tcp_payload = tcp && doff >= 20 && po < end;
        if (fip == coredns && fport == 53) { if (udp) nm = NM_DNSQ; else if (tcp_payload) nm = NM_DNSQ_TCP; }
        else if (tcp_payload && fip == envoy && fport == egress_port) nm = NM_TLS;
    }
    if (metric) { hm = true; mk = ((unsigned long long)((mp & 0xffffu) | (crow << 16) | (metric << 30)) << 32) | mh; }   // == metric_pack
}

template <bool kName>
__device__ __forceinline__ bool fast_classify(const DevTables &T, const Tabs &tb, uint32_t w3, const uint4 &q1, uint32_t q2x, uint32_t q2y,
                                              uint32_t q2w, uint32_t avail, uint32_t flags, uint32_t &word, bool &hm, unsigned long long &mk, uint32_t &nm)
{
    nm = NM_NONE;
    // Ethernet II + IPv4 without options, not a later fragment
    if ((w3 & 0x00ffffffu) != 0x00450008u || (q1.y & 0xff1fu) != 0) {
        const uint32_t et = w3 & 0xffffu;                                        // ethertype, network order in the low half
        if (et != 0x0008u && et != 0xDD86u) { word = 0; hm = false; mk = 0; nm = NM_NONE; return true; }   // not IP: no socket event
        return false;
    }
    const uint32_t proto = q1.y >> 24;
    const uint32_t src = __byte_perm(q1.z, q1.w, 0x5432), dst = __byte_perm(q1.w, q2x, 0x5432);
    const uint32_t sport = bswap16(q2x >> 16), dport = bswap16(q2y);
    uint32_t end = 0, doff = 0;
    if (kName) {                                                                 // TCP payload window from the header words already in registers
        end = 14 + bswap16(q1.x); if (end > avail) end = avail;
        doff = ((q2w >> 20) & 0xfu) * 4;                                         // q2w = bytes 44..47: the data offset is byte 46
    }
    fast_fields<kName>(T, tb, proto, src, dst, sport, dport, end, doff, flags, word, hm, mk, nm);
    return true;
}

// Every warp runs its own pipeline: a private ring of kStages 2 KiB tiles (32 slots x 64 B, filled by
// TMA and signalled through the warp's own mbarriers), so there is no CTA-wide barrier in the loop.
// The tile is copied to registers as soon as it lands and refilled immediately, which is why one
// stage per warp already keeps the copy engine ahead of the ALU work.  Rare packets go to a per-warp
// queue that is classified 32 at a time with all lanes converged.
template <int kStages, bool kSwizzle, int kMinBlocks, bool kAllStaged>
__global__ void __launch_bounds__(FAST_WARPS * 32, kMinBlocks) k_slots_fast(const __grid_constant__ DevTables T, const __grid_constant__ CUtensorMap map,
                                                                const uint8_t *__restrict__ pkts, uint32_t stride,
                                                                uint32_t *__restrict__ rare_q, uint32_t *__restrict__ rare_cnt, uint32_t agg_slots,
                                                                unsigned long long n, uint32_t flags, uint32_t *__restrict__ verdicts)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemLayout L = smem_layout(T, FAST_WARPS * kStages * WTILE_BYTES, FAST_WARPS * WQ_CAP * 4, FAST_WARPS * kStages, agg_slots, FAST_WARPS);
    Tabs tb; MetricAgg ag;
    cta_prologue<kAllStaged>(T, smem, L, FAST_WARPS * kStages, tb, ag, agg_slots);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = UNIFORM_WARP_IDX();
    uint64_t *bars = (uint64_t *)(smem + L.bars) + warp * kStages;
    uint8_t *tiles = smem + L.tiles + warp * (kStages * WTILE_BYTES);
    uint32_t *queue = (uint32_t *)(smem + L.queue) + warp * WQ_CAP;
    const uint32_t n_tiles = (uint32_t)((n + 31ull) >> 5);   // n < 2^31 on this path
    const uint32_t gw = blockIdx.x * FAST_WARPS + warp, W = gridDim.x * FAST_WARPS;
    const bool do_metrics = !(flags & CLK_F_NO_METRICS);
    __syncthreads();                                         // tables staged, aggregation table zeroed, barriers initialised

    if (elect_one()) {
        for (int s = 0; s < kStages; s++) {
            uint32_t t = gw + (uint32_t)s * W;
            if (t < n_tiles) { mbar_expect_tx(&bars[s], WTILE_BYTES); tma_load_2d(tiles + s * WTILE_BYTES, &map, &bars[s], 0, (int)(t << 5)); }
        }
    }
    // rare frames: per-warp staging queue -> this CTA's own region of rare_q (no cross-CTA traffic, no grid barrier)
    const uint32_t cta_cap = ((n_tiles + W - 1) / W) * FAST_WARPS * 32;
    uint32_t *cta_q = rare_q + (size_t)blockIdx.x * cta_cap;
    uint32_t *cta_cnt = (uint32_t *)(smem + L.bars) + 2 * (FAST_WARPS * kStages + 1);   // one word after the mbarriers
    uint32_t qn = 0;                                         // entries in this warp's queue (warp-uniform)
    uint32_t it = 0;
    for (uint32_t t = gw; t < n_tiles; t += W, it++) {
        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
        mbar_wait(&bars[s], ph);
        const uint8_t *slot = tiles + s * WTILE_BYTES + lane * 64;
        const uint32_t x = kSwizzle ? ((lane >> 1) & 3u) : 0u;               // SWIZZLE_64B: chunk ^= (row >> 1) & 3
        const uint32_t w3 = *(const uint32_t *)(slot + ((0u ^ x) << 4) + 12);
        const uint4 q1 = *(const uint4 *)(slot + ((1u ^ x) << 4));
        const uint2 q2 = *(const uint2 *)(slot + ((2u ^ x) << 4));
        __syncwarp();                                                        // every lane holds its header words
        if (elect_one()) {
            uint32_t tn = t + (uint32_t)kStages * W;
            // The fence is load-bearing: it makes the elected lane wait until the warp's ld.shared of this tile have RETURNED
            // before the async proxy may overwrite the stage (__syncwarp alone only orders their issue).  Measured without it
            // (round 2): verdict words of the checked copies still equal, but 1-2 stray metrics_map rows per 10 M packets —
            // a header read after the refill had landed — for +0.3 %.
            if (tn < n_tiles) { fence_proxy_async(); mbar_expect_tx(&bars[s], WTILE_BYTES); tma_load_2d(tiles + s * WTILE_BYTES, &map, &bars[s], 0, (int)(tn << 5)); }
        }
        const uint32_t i = (t << 5) + lane;
        bool hm = false, slow = false; unsigned long long mk = 0; uint32_t w = 0;
        if (i < n) {
            uint32_t nm; slow = !fast_classify<false>(T, tb, w3, q1, q2.x, q2.y, 0, 64, 0, w, hm, mk, nm);
            if (!slow) stg_stream(&verdicts[i], w);
        }
        const unsigned sm = __ballot_sync(0xffffffffu, slow);
        if (sm) {
            if (slow) queue[qn + __popc(sm & ((1u << lane) - 1u))] = i;
            qn += __popc(sm);
            __syncwarp();
        }
        if (do_metrics) metric_warp_emit(T, ag, hm, mk);
        if (qn >= 32) {                                                      // move 32 rare frames to this CTA's region of rare_q
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cta_cnt, 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            cta_q[base + lane] = queue[lane];
            const uint32_t carry = lane + 32 < qn ? queue[lane + 32] : 0;
            __syncwarp();
            if (lane + 32 < qn) queue[lane] = carry;
            qn -= 32;
            __syncwarp();
        }
    }
    if (qn) {                                                                // leftovers (< 32)
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cta_cnt, qn);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < qn) cta_q[base + lane] = queue[lane];
    }
    // ---- rare frames (IPv6, IPv4 options, malformed; a few % of a realistic mix) ------------------------------
    // Each CTA classifies the rare frames of its own tiles once its main loop is done: the generic path with every
    // lane busy, tables and the metric cache still in shared memory.
    __threadfence_block();
    __syncthreads();
    {
        const uint32_t cnt = *(volatile uint32_t *)cta_cnt, cnt_round = (cnt + 31u) & ~31u;
        for (uint32_t k = tid; k < cnt_round; k += FAST_WARPS * 32) {
            bool hm = false; unsigned long long mk = 0;
            if (k < cnt) {
                const unsigned long long j = cta_q[k];
                stg_stream(&verdicts[j], generic_one<false>(T, tb, pkts + j * stride, stride, flags & ~CLK_F_NAME_STAGE, hm, mk));
            }
            if (do_metrics) metric_warp_emit(T, ag, hm, mk);
        }
    }
    if (do_metrics) metric_cta_flush(T, ag);
}

// ------------------------------------------------------------------------------------------------
// kernels 1b: variable-length captures with the name stage (BASELINE config 3, the north-star workload)
// ------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ void payload_window(P pk, uint32_t avail, bool tcp, uint32_t &po, uint32_t &pl)
{
    // simple IPv4 (ihl 5): l4off = 34; same arithmetic as name_stage()
    uint32_t end = 14 + gbe16(pk + 16u); if (end > avail) end = avail;
    po = 0; pl = 0;
    if (!tcp) { po = 42; if (po <= end) pl = end - po; }
    else if (34 + 13 <= avail) { uint32_t doff = (gb(pk + 46u) >> 4) * 4; po = 34 + doff; if (doff >= 20 && po <= end) pl = end - po; }
}

// Measured on B200 (profiles/r02_varlen_notes.md): a fused kernel that holds the header ladder and every payload
// parser executes ~100 KB of SASS from warps in different phases — far more than the 32 KB L1.5 instruction cache —
// and stalls on instruction fetch (no_instruction 53 % of warp stalls); a CTA that stages chunks in shared memory and
// runs the parsers between barriers pays every long-latency event of any warp at the next barrier.  So the work is cut
// BY ROLE into kernels whose code is small and whose warps all run the same thing:
//   k_varlen_head     every packet: the header ladder straight from global memory (three 16-byte loads per packet, no
//                     CTA barrier in the loop), verdicts of packets that need nothing else, metrics; packets that need
//                     payload work are appended — 32 at a time per warp and type, through shared-memory queues — to one
//                     global list PER WORK TYPE {index, phase-A word, frame offset, frame length};
//   k_name_list<type> one parser per launch (DNS question / DNS-over-TCP / ClientHello / DNS answer / HTTP request):
//                     a warp takes 32 list entries, so its lanes are full and run the same code in lock step;
//   k_generic_list    the frames the fast ladder does not cover (IPv6, IP options, runts): the generic path.
// A batch is cut into slices so that the lists stay bounded (VL_SLICE packets); everything is enqueued on one stream.
#define VW_LISTS 6             // 0 generic, then NM_DNSQ .. NM_HTTP
#define VW_QCAP 32             // a warp's queue is flushed (wholly) before it would overflow, and when it is full

// a work-list entry is one uint4: {packet index, phase-A verdict word, frame offset / 16, frame length / 16}

template <int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks) k_varlen_head(const __grid_constant__ DevTables T, const uint8_t *__restrict__ pkts,
                                                                         const uint32_t *__restrict__ offs16, uint32_t first, uint32_t n,
                                                                         uint32_t flags, uint32_t *__restrict__ verdicts,
                                                                         uint4 *__restrict__ ents, uint32_t *__restrict__ cnt, uint32_t cap)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemLayout L = smem_layout(T, 0, kWarps * VW_LISTS * VW_QCAP * 16, 0, MSLOTS, kWarps);
    Tabs tb; MetricAgg ag;
    cta_prologue<false>(T, smem, L, 0, tb, ag);
    __syncthreads();
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = UNIFORM_WARP_IDX();
    uint4 *queue = (uint4 *)(smem + L.queue) + warp * (VW_LISTS * VW_QCAP);
    const uint32_t n_tiles = (n + 31u) >> 5;
    const uint32_t gw = blockIdx.x * kWarps + warp, W = gridDim.x * kWarps;
    const bool do_metrics = !(flags & CLK_F_NO_METRICS);
    uint32_t qn[VW_LISTS];
    #pragma unroll
    for (int k = 0; k < VW_LISTS; k++) qn[k] = 0;

    // the whole queue k of this warp -> global list k: one atomic, one coalesced store of 16-byte entries
    auto flush = [&](int k) {
        const uint32_t take = qn[k];
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&cnt[k], take);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < take && base + lane < cap) ents[(size_t)k * cap + base + lane] = queue[k * VW_QCAP + lane];
        qn[k] = 0;
        __syncwarp();
    };

    // Software pipeline, two tiles deep: while tile t is classified, the header words of tile t + W and the offsets of
    // tile t + 2W are in flight (a warp keeps ~3 KB of loads outstanding; 16 warps cover the DRAM latency of an SM).
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    uint32_t t = gw;
    uint32_t o0 = 0, o1 = 0, o0n = 0, o1n = 0, w3 = 0; uint4 q1 = z4, q2 = z4;
    if (t < n_tiles) { const uint32_t i = (t << 5) + lane; if (i < n) { o0 = __ldg(&offs16[first + i]); o1 = __ldg(&offs16[first + i + 1]); } }
    if (t + W < n_tiles) { const uint32_t i = ((t + W) << 5) + lane; if (i < n) { o0n = __ldg(&offs16[first + i]); o1n = __ldg(&offs16[first + i + 1]); } }
    if (t < n_tiles && (o1 - o0) >= 4u) {
        const uint8_t *pk = pkts + (unsigned long long)o0 * 16ull;
        w3 = __ldg((const uint32_t *)(pk + 12)); q1 = __ldg((const uint4 *)(pk + 16)); q2 = __ldg((const uint4 *)(pk + 32));
    }
    for (; t < n_tiles; t += W) {
        const uint32_t i = (t << 5) + lane;
        // next tile's header words, the tile after's offsets
        uint32_t w3n = 0, o0f = 0, o1f = 0; uint4 q1n = z4, q2n = z4;
        if (t + W < n_tiles && (o1n - o0n) >= 4u) {
            const uint8_t *pk = pkts + (unsigned long long)o0n * 16ull;
            w3n = __ldg((const uint32_t *)(pk + 12)); q1n = __ldg((const uint4 *)(pk + 16)); q2n = __ldg((const uint4 *)(pk + 32));
        }
        if (t + 2 * W < n_tiles) { const uint32_t i2 = ((t + 2 * W) << 5) + lane; if (i2 < n) { o0f = __ldg(&offs16[first + i2]); o1f = __ldg(&offs16[first + i2 + 1]); } }

        bool hm = false; unsigned long long mk = 0; uint32_t w = 0, kind = 0xffu;   // kind: list to join, 0xff = done here
        if (i < n) {
            const uint32_t avail = (o1 - o0) * 16u;
            kind = 0;
            if (avail >= 64) {
                uint32_t nm;
                if (fast_classify<true>(T, tb, w3, q1, q2.x, q2.y, q2.w, avail, flags, w, hm, mk, nm)) kind = nm == NM_NONE ? 0xffu : nm;
            }
            if (kind == NM_DNSQ_TCP) { kind = 0; hm = false; }   // DNS over TCP is rare: the generic path redoes the packet (verdict, metric and name),
                                                               // one list and one launch fewer — so nothing is counted for it here
            if (kind == 0xffu) stg_stream(&verdicts[first + i], w);
        }
        if (do_metrics) metric_warp_emit(T, ag, hm, mk);
        #pragma unroll
        for (int k = 0; k < VW_LISTS; k++) {
            if (k == (int)NM_HTTP || k == (int)NM_DNSQ_TCP) continue;           // HTTP is filled by the ClientHello kernel, not here
            const unsigned m = __ballot_sync(0xffffffffu, kind == (uint32_t)k);
            if (m) {
                const uint32_t c = __popc(m);
                if (qn[k] + c > VW_QCAP) flush(k);
                if (kind == (uint32_t)k) queue[k * VW_QCAP + qn[k] + __popc(m & ((1u << lane) - 1u))] = make_uint4(first + i, w, o0, o1 - o0);
                qn[k] += c;
                __syncwarp();
                if (qn[k] == VW_QCAP) flush(k);
            }
        }
        o0 = o0n; o1 = o1n; w3 = w3n; q1 = q1n; q2 = q2n; o0n = o0f; o1n = o1f;
    }
    #pragma unroll
    for (int k = 0; k < VW_LISTS; k++) if (qn[k]) flush(k);
    if (do_metrics) metric_cta_flush(T, ag);
}

// One parser per instantiation.  Every lane of `mask` holds a packet of work type kType and calls together.
template <int kType>
__device__ __forceinline__ uint32_t name_work_t(const DevTables &T, const uint8_t *pk, uint32_t avail, uint32_t word, unsigned mask)
{
    uint32_t po, pl;
    if (kType == (int)NM_DNSA) {
        uint32_t end = 14 + gbe16(pk + 16u); if (end > avail) end = avail;
        const bool has = 42 <= end;
        uint32_t nv, nupd = dns_answer_updates(T, pk + 42u, has ? end - 42 : 0u, 0, nullptr, 0, 0, &nv, mask);
        return has ? (nv << 4) | (nupd << 27) : 0u;
    }
    payload_window(pk, avail, kType != (int)NM_DNSQ, po, pl);
    uint32_t name = CLK_N_NA;
    if (kType == (int)NM_DNSQ) name = dns_query_verdict(T, pk + po, pl, mask);
    else if (kType == (int)NM_DNSQ_TCP) {
        const bool framed = pl >= 2 && gbe16(pk + po) + 2 <= pl;
        const uint32_t v = dns_query_verdict(T, pk + (po + 2), framed ? gbe16(pk + po) : 0u, mask);
        if (pl > 0) name = framed ? v : CLK_N_DENY;
    }
    else if (kType == (int)NM_HTTP) { if (pl > 0) name = http_host_verdict(T, pk + po, pl); }
    else { const uint32_t v = tls_sni_verdict(T, pk + po, pl, mask); if (pl > 0) name = v; }
    return word | (name << 4);
}

template <int kType, int kMinB>
__global__ void __launch_bounds__(256, kMinB) k_name_list(const __grid_constant__ DevTables T, const uint8_t *__restrict__ pkts, uint4 *__restrict__ ents,
                                                   uint32_t *__restrict__ cnt, uint32_t cap, uint32_t *__restrict__ verdicts)
{
    const uint32_t n_k = cnt[kType] < cap ? cnt[kType] : cap;
    const uint32_t lane = threadIdx.x & 31u, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, W = (gridDim.x * blockDim.x) >> 5;
    const uint4 *mine = ents + (size_t)kType * cap;
    // The parsers walk their frame through chains of dependent byte loads; a frame's lines beyond the first 64 bytes have not
    // been touched since the capture was written, so every new line would cost a DRAM round trip in the middle of a chain.
    // Each lane therefore asks for ALL lines of its frame at once (L2 prefetch, no register cost) — for this unit right away
    // and for the warp's next unit one unit ahead.
    auto prefetch_frame = [&](const uint4 &en) {
        const uint8_t *q = pkts + (unsigned long long)en.z * 16ull; const uint32_t bytes = en.w * 16u;
        for (uint32_t o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
    };
    uint4 ent_n = make_uint4(0, 0, 0, 0);
    if (gw * 32u + lane < n_k) { ent_n = mine[gw * 32u + lane]; prefetch_frame(ent_n); }
    for (uint32_t u = gw; u * 32u < n_k; u += W) {
        const uint32_t e = u * 32u + lane; const bool valid = e < n_k;
        const uint4 ent = ent_n;
        { const uint32_t en = (u + W) * 32u + lane; ent_n = make_uint4(0, 0, 0, 0); if (en < n_k) { ent_n = mine[en]; prefetch_frame(ent_n); } }
        const uint8_t *pk = pkts + (unsigned long long)ent.z * 16ull; const uint32_t avail = ent.w * 16u;
        bool defer = false;
        if (kType == (int)NM_TLS && T.n_http) {                 // plaintext HTTP shares the egress port with TLS: those go to the HTTP list
            if (valid) { uint32_t po, pl; payload_window(pk, avail, true, po, pl); defer = pl >= 1 && gb(pk + po) != 0x16u; }
            const unsigned dm = __ballot_sync(0xffffffffu, defer);
            if (dm) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&cnt[NM_HTTP], (uint32_t)__popc(dm));
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t pos = base + __popc(dm & ((1u << lane) - 1u));
                if (defer && pos < cap) ents[(size_t)NM_HTTP * cap + pos] = ent;
            }
        }
        const bool run = valid && !defer;
        const unsigned um = __ballot_sync(0xffffffffu, run);
        if (run) stg_stream(&verdicts[ent.x], name_work_t<kType>(T, pk, avail, ent.y, um));
    }
}

// List 0: the frames the fast ladder does not cover — the generic path (its metrics included).
__global__ void __launch_bounds__(256) k_generic_list(const __grid_constant__ DevTables T, const uint8_t *__restrict__ pkts, const uint4 *__restrict__ ents,
                                                      const uint32_t *__restrict__ cnt, uint32_t cap, uint32_t flags, uint32_t *__restrict__ verdicts)
{
    const uint32_t n_k = cnt[0] < cap ? cnt[0] : cap;
    if (n_k == 0) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    Tabs tb; MetricAgg ag;
    cta_prologue<false>(T, smem, L, 0, tb, ag);
    __syncthreads();
    const bool do_metrics = !(flags & CLK_F_NO_METRICS);
    const uint32_t step = gridDim.x * blockDim.x, n_round = (n_k + 31u) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += step) {
        bool hm = false; unsigned long long mk = 0;
        if (k < n_k) {
            const uint4 ent = ents[k];
            stg_stream(&verdicts[ent.x], generic_one<true>(T, tb, pkts + (unsigned long long)ent.z * 16ull, ent.w * 16u, flags, hm, mk));
        }
        if (do_metrics) metric_warp_emit(T, ag, hm, mk);
    }
    if (do_metrics) metric_cta_flush(T, ag);
}

// ------------------------------------------------------------------------------------------------
// kernel 2: direct loads (any stride, or varlen offsets) + optional name stage (BASELINE config 3)
// ------------------------------------------------------------------------------------------------
template <bool kVarlen>
__global__ void __launch_bounds__(256) k_generic(const __grid_constant__ DevTables T, const uint8_t *__restrict__ pkts,
                                                 const uint32_t *__restrict__ offs16, uint32_t stride, unsigned long long n,
                                                 uint32_t flags, uint32_t *__restrict__ verdicts)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    Tabs tb; MetricAgg ag;
    cta_prologue<false>(T, smem, L, 0, tb, ag);
    __syncthreads();
    const bool do_metrics = !(flags & CLK_F_NO_METRICS);
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_round = (n + 31ull) & ~31ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
        bool hm = false; unsigned long long mk = 0;
        if (i < n) {
            const uint8_t *pk; uint32_t avail;
            if (kVarlen) { uint32_t o0 = __ldg(&offs16[i]), o1 = __ldg(&offs16[i + 1]); pk = pkts + (unsigned long long)o0 * 16ull; avail = (o1 - o0) * 16u; }
            else { pk = pkts + i * stride; avail = stride; }
            stg_stream(&verdicts[i], generic_one<true>(T, tb, pk, avail, flags, hm, mk));
        }
        if (do_metrics) metric_warp_emit(T, ag, hm, mk);
    }
    if (do_metrics) metric_cta_flush(T, ag);
}

// ------------------------------------------------------------------------------------------------
// kernel 2b: compact ingest records (clk_hdr16): 16 bytes per frame instead of the frame
// ------------------------------------------------------------------------------------------------
// What the decision reads of an IPv4 header-only frame fits 16 bytes; a capture front end (clk_pack_hdr16, or an
// AF_PACKET / XDP ring reader) emits one record per frame, so the host link carries 16 B per packet instead of 64.
// Frames the record cannot express (IPv6, IP options, truncated) travel whole in a side array of 64-byte slots.
__global__ void __launch_bounds__(256) k_hdr16(const __grid_constant__ DevTables T, const uint4 *__restrict__ recs, const uint8_t *__restrict__ aux,
                                               unsigned long long n, uint32_t flags, uint32_t *__restrict__ verdicts)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    Tabs tb; MetricAgg ag;
    cta_prologue<false>(T, smem, L, 0, tb, ag);
    __syncthreads();
    const bool do_metrics = !(flags & CLK_F_NO_METRICS);
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_round = (n + 31ull) & ~31ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
        bool hm = false; unsigned long long mk = 0;
        if (i < n) {
            const uint4 r = __ldg(&recs[i]);
            const uint32_t kind = (r.w >> 8) & 0xffu;
            uint32_t w = 0, nm;
            if (kind == CLK_HDR_IPV4) fast_fields<false>(T, tb, r.w & 0xffu, r.x, r.y, r.z & 0xffffu, r.z >> 16, 0, 0, 0, w, hm, mk, nm);
            else if (kind == CLK_HDR_AUX) w = generic_one<false>(T, tb, aux + (unsigned long long)r.x * 64ull, 64, flags & ~CLK_F_NAME_STAGE, hm, mk);
            stg_stream(&verdicts[i], w);
        }
        if (do_metrics) metric_warp_emit(T, ag, hm, mk);
    }
    if (do_metrics) metric_cta_flush(T, ag);
}

// ------------------------------------------------------------------------------------------------
// kernel 3: pre-parsed socket events (the reference hooks' own input)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_events(const __grid_constant__ DevTables T, const clk_event *__restrict__ ev,
                                                unsigned long long n, uint32_t flags, uint32_t *__restrict__ verdicts)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    Tabs tb; MetricAgg ag;
    cta_prologue<false>(T, smem, L, 0, tb, ag);
    __syncthreads();
    const bool do_metrics = !(flags & CLK_F_NO_METRICS);
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_round = (n + 31ull) & ~31ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
        bool hm = false; unsigned long long mk = 0;
        if (i < n) {
            clk_event e = ev[i];
            Decision d; d.verdict = 0; d.metric = 0; d.new_ip = 0; d.new_port = 0; d.m_hash = 0; d.m_port = 0;
            // container_map / bypass_map by cgroup id
            const ContRow *cr = nullptr; uint32_t crow = 0;
            if (T.n_crows) {
                uint32_t h = mix64(e.cgroup_id) & T.cg_mask;
                for (;;) { uint32_t s = T.cg_slots[h]; if (!s) break; if (T.crows[s - 1].cgroup_id == e.cgroup_id) { cr = &T.crows[s - 1]; crow = s - 1; break; } h = (h + 1) & T.cg_mask; }
            }
            const bool six = e.hook >= CLK_HOOK_CONNECT6 && e.hook <= CLK_HOOK_RECVMSG6;
            const uint32_t ip = six ? e.user_ip6[3] : e.user_ip4;
            uint32_t w = 0;
            if (cr) {
                run_hooks(T, tb, e.hook, e.sock_type, e.uid, e.so_mark, cr->flags, &cr->cfg, ip, e.user_port,
                          e.user_ip6[0], e.user_ip6[1], e.user_ip6[2], d);
                if (d.verdict == CLK_V_REWRITE && d.new_ip == ip && d.new_port == e.user_port) d.verdict = CLK_V_PASS;
                if (d.metric) { hm = true; mk = metric_pack(d.m_hash, d.m_port, crow, d.metric - 1); }
                w = make_word(d, &cr->cfg, (cr->flags & CLK_ROW_ENFORCED) != 0, e.hook, 0, 0);
            } else {
                w = (uint32_t)e.hook << 24;             // cgroup in neither map: every program returns 1 untouched
            }
            stg_stream(&verdicts[i], w);
        }
        if (do_metrics) metric_warp_emit(T, ag, hm, mk);
    }
    if (do_metrics) metric_cta_flush(T, ag);
}

// ------------------------------------------------------------------------------------------------
// kernel 4: dnsbpf update extraction (second pass over DNS-answer packets only)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dns_extract(const __grid_constant__ DevTables T, const uint8_t *__restrict__ pkts,
                                                     const uint32_t *__restrict__ offs16, unsigned long long n,
                                                     const uint32_t *__restrict__ verdicts, const unsigned long long *__restrict__ out_pos,
                                                     clk_dns_update_rec *__restrict__ out, unsigned long long cap)
{
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        uint32_t w = verdicts[i];
        if (CLK_HOOK(w) != CLK_HOOK_NONE || CLK_NUPD(w) == 0) continue;
        uint32_t o0 = offs16[i], o1 = offs16[i + 1];
        dns_answer_extract(T, pkts + (unsigned long long)o0 * 16ull, (o1 - o0) * 16u, (uint32_t)i, out, out_pos[i], cap);
    }
}

// per-packet update counts -> exclusive prefix positions (counts are tiny and the array is streamed
// once; a block-sum / serial-scan / per-block-scan trio keeps it simple and deterministic)
__global__ void __launch_bounds__(256) k_count_block(const uint32_t *__restrict__ verdicts, unsigned long long n, unsigned long long *__restrict__ block_sums)
{
    __shared__ unsigned long long sh[8];
    unsigned long long base = (unsigned long long)blockIdx.x * 4096ull, s = 0;
    for (uint32_t k = threadIdx.x; k < 4096; k += 256) { unsigned long long i = base + k; if (i < n) { uint32_t w = verdicts[i]; if (CLK_HOOK(w) == CLK_HOOK_NONE) s += CLK_NUPD(w); } }
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int i = 0; i < 8; i++) t += sh[i]; block_sums[blockIdx.x] = t; }
}
__global__ void k_scan_blocks(unsigned long long *block_sums, uint32_t nb, unsigned long long *total)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long run = 0; for (uint32_t i = 0; i < nb; i++) { unsigned long long v = block_sums[i]; block_sums[i] = run; run += v; } *total = run; }
}
__global__ void __launch_bounds__(256) k_positions(const uint32_t *__restrict__ verdicts, unsigned long long n,
                                                   const unsigned long long *__restrict__ block_sums, unsigned long long *__restrict__ out_pos)
{
    __shared__ unsigned long long carry;
    unsigned long long base = (unsigned long long)blockIdx.x * 4096ull;
    if (threadIdx.x == 0) carry = block_sums[blockIdx.x];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    unsigned long long run = carry;
    for (uint32_t k = 0; k < 4096; k += 32) {
        unsigned long long i = base + k + threadIdx.x; uint32_t c = 0;
        if (i < n) { uint32_t w = verdicts[i]; if (CLK_HOOK(w) == CLK_HOOK_NONE) c = CLK_NUPD(w); }
        uint32_t inc = c;
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if ((int)threadIdx.x >= o) inc += v; }
        if (i < n) out_pos[i] = run + inc - c;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// ------------------------------------------------------------------------------------------------
// counters block export / import (the payload of the one multi-GPU collective)
// ------------------------------------------------------------------------------------------------
// Compacted export: block[0] is a header whose `count` is the number of rows that follow; rows 1..count are
// the occupied metrics_map entries in no particular order (cap_rows bounds the block).
__global__ void k_counters_export(const __grid_constant__ DevTables T, clk_metric *__restrict__ block, uint32_t cap_rows)
{
    const uint32_t lane = threadIdx.x & 31u, slots = T.m_mask + 1, slots_round = (slots + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slots_round; i += gridDim.x * blockDim.x) {
        const unsigned long long k = i < slots ? T.m_keys[i] : 0ull;
        const unsigned m = __ballot_sync(0xffffffffu, k != 0ull);
        if (!m) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd((unsigned long long *)&block[0].count, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (k) {
            const unsigned long long pos = 1 + base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap_rows) {
                clk_metric r; const uint32_t crow = (uint32_t)(k >> 48) & 0x3fffu;
                r.key.cgroup_id = T.crows[crow].cgroup_id; r.key.domain_hash = (uint32_t)k; r.key.dst_port = (uint16_t)(k >> 32);
                r.key.action = (uint8_t)((k >> 62) - 1); r.key._pad = 0; r.count = T.m_counts[i];
                block[pos] = r;
            }
        }
    }
}
// Dense per-container totals of the local metrics_map: block[r] = {cgroup_id of container row r, ALLOW, DENY, BYPASS}.
// This is the small per-batch collective payload (rows x 32 B); the caller zeroes `block` first.  A warp groups its 32
// slots by (container row, action) and sums each group with three REDUX adds over 22-bit pieces of the counts.
__global__ void k_counters_totals(const __grid_constant__ DevTables T, clk_total *__restrict__ block, uint32_t rows)
{
    extern __shared__ unsigned long long tot[];                            // n_crows x 3 when it fits
    const uint32_t lane = threadIdx.x & 31u, slots = T.m_mask + 1, slots_round = (slots + 31u) & ~31u;
    const uint32_t nr = T.n_crows < rows ? T.n_crows : rows;
    const bool use_smem = nr * 24u <= 48u * 1024u;
    if (use_smem) { for (uint32_t i = threadIdx.x; i < nr * 3u; i += blockDim.x) tot[i] = 0ull; __syncthreads(); }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slots_round; i += gridDim.x * blockDim.x) {
        const unsigned long long k = i < slots ? T.m_keys[i] : 0ull;
        const unsigned long long c = k ? T.m_counts[i] : 0ull;
        const uint32_t crow = (uint32_t)(k >> 48) & 0x3fffu, act = (uint32_t)(k >> 62) - 1u;
        const bool has = k != 0ull && c != 0ull && crow < nr && act < 3u;
        const uint32_t idx = has ? crow * 3u + act : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, idx);
        const uint32_t s0 = __reduce_add_sync(peers, (uint32_t)c & 0x3fffffu);
        const uint32_t s1 = __reduce_add_sync(peers, (uint32_t)(c >> 22) & 0x3fffffu);
        const uint32_t s2 = __reduce_add_sync(peers, (uint32_t)(c >> 44));
        if (has && (int)lane == __ffs(peers) - 1) {
            const unsigned long long sum = (unsigned long long)s0 + ((unsigned long long)s1 << 22) + ((unsigned long long)s2 << 44);
            if (use_smem) atomicAdd(&tot[idx], sum); else atomicAdd((unsigned long long *)&block[crow].count[act], sum);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nr * 3u; i += blockDim.x) if (tot[i]) atomicAdd((unsigned long long *)&block[i / 3u].count[i % 3u], tot[i]);
    }
    if (blockIdx.x == 0) for (uint32_t r = threadIdx.x; r < nr; r += blockDim.x) block[r].cgroup_id = T.crows[r].cgroup_id;
}
__global__ void k_counters_clear(const __grid_constant__ DevTables T)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.m_mask; i += gridDim.x * blockDim.x) { T.m_keys[i] = 0ull; T.m_counts[i] = 0ull; }
    if (blockIdx.x == 0 && threadIdx.x == 0) { T.m_dropped[0] = 0ull; T.m_dropped[1] = 0ull; }
}
__global__ void k_counters_import(const __grid_constant__ DevTables T, const clk_metric *__restrict__ blocks, uint32_t n_blocks, uint32_t block_rows)
{
    const unsigned long long total = (unsigned long long)n_blocks * block_rows;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(i / block_rows), r = (uint32_t)(i % block_rows);
        if (r == 0) continue;                                               // header
        unsigned long long valid = blocks[(unsigned long long)b * block_rows].count;
        if (valid > block_rows - 1) valid = block_rows - 1;
        if (r > valid) continue;
        clk_metric m = blocks[i];
        if (m.count == 0) continue;
        uint32_t h = mix64(m.key.cgroup_id) & T.cg_mask, crow = 0xffffffffu;
        if (T.n_crows) for (;;) { uint32_t s = T.cg_slots[h]; if (!s) break; if (T.crows[s - 1].cgroup_id == m.key.cgroup_id) { crow = s - 1; break; } h = (h + 1) & T.cg_mask; }
        if (crow == 0xffffffffu) { atomicAdd(&T.m_dropped[0], m.count); continue; }
        {   // counts can exceed 32 bits after merging: add in <= 2^31 pieces
            unsigned long long left = m.count; const unsigned long long k = metric_pack(m.key.domain_hash, m.key.dst_port, crow, m.key.action);
            while (left) { unsigned int piece = left > 0x7fffffffull ? 0x7fffffffu : (unsigned int)left; unsigned int dr = metric_global_add(T, k, piece); if (dr) atomicAdd(&T.m_dropped[0], (unsigned long long)dr); left -= piece; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count()
{
    if (!g_sm_count) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev); if (g_sm_count <= 0) g_sm_count = 148; }
    return g_sm_count;
}

// SMs the persistent classify kernels leave free (clk_cfg.reserved_sms), so that the counter totals / export /
// NCCL all-gather / import of batch k (side stream) run beside the classify kernel of batch k+1 instead of
// delaying its CTAs: a classify CTA owns a whole SM (64 K registers), nothing can co-reside with it.
static int reserved_sms(const DevTables &T) { int cap = sm_count() / 2, r = (int)T.reserved_sms; return r > cap ? cap : r; }
static int side_grid(const DevTables &T, int want) { int r = reserved_sms(T); return r > 0 ? (r * 8 < want ? r * 8 : want) : want; }   // 256-thread CTAs, 8 per SM

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode()
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}

template <typename K>
static cudaError_t occupancy_grid(const DevTables &T, K kernel, int threads, size_t smem, int *grid)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *grid = per_sm * (sm_count() - reserved_sms(T));              // persistent kernels leave the reserved SMs to the collective
    return cudaSuccess;
}

template <int kStages, bool kSwizzle, int kMinBlocks, bool kAllStaged>
static const char *launch_fast(const DevTables &T, const CUtensorMap &map, const void *d_pkts, uint32_t stride, unsigned long long n,
                               uint32_t flags, uint32_t *d_verdicts, cudaStream_t st, uint32_t *rare_q, uint32_t *rare_cnt, LaunchInfo *info)
{
    int grid = 0; cudaError_t e;
    // metric cache: the largest power of two that still fits kMinBlocks CTAs per SM (227 KiB per SM, 1 KiB reserved per CTA)
    uint32_t agg_slots = 512;
    const uint32_t budget = (227u * 1024u) / kMinBlocks - 1024u;
    while (agg_slots < 2048 && smem_layout(T, FAST_WARPS * kStages * WTILE_BYTES, FAST_WARPS * WQ_CAP * 4, FAST_WARPS * kStages, agg_slots * 2, FAST_WARPS).total <= budget) agg_slots *= 2;
    if (const char *s = getenv("CLK_AGG_SLOTS")) { uint32_t v = (uint32_t)atoi(s); if (v >= 64 && (v & (v - 1)) == 0) agg_slots = v; }
    SmemLayout L = smem_layout(T, FAST_WARPS * kStages * WTILE_BYTES, FAST_WARPS * WQ_CAP * 4, FAST_WARPS * kStages, agg_slots, FAST_WARPS);
    if ((e = occupancy_grid(T, k_slots_fast<kStages, kSwizzle, kMinBlocks, kAllStaged>, FAST_WARPS * 32, L.total, &grid)) != cudaSuccess) return cudaGetErrorString(e);
    unsigned long long n_cta = ((n + 31) / 32 + FAST_WARPS - 1) / FAST_WARPS;
    if ((unsigned long long)grid > n_cta) grid = (int)n_cta;
    k_slots_fast<kStages, kSwizzle, kMinBlocks, kAllStaged><<<grid, FAST_WARPS * 32, L.total, st>>>(T, map, (const uint8_t *)d_pkts, stride, rare_q, rare_cnt, agg_slots, n, flags, d_verdicts);
    if (info) { info->grid = grid; info->block = FAST_WARPS * 32; info->smem = L.total; }
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char *launch_slots(const DevTables &T, const void *d_pkts, uint32_t stride, unsigned long long n, uint32_t flags,
                         uint32_t *d_verdicts, cudaStream_t st, int loader, int stages, uint32_t *rare_q, uint32_t *rare_cnt, LaunchInfo *info)
{
    if (n == 0) return nullptr;
    cudaError_t e;
    int grid = 0;
    if (loader == LOADER_DIRECT || (flags & CLK_F_NAME_STAGE) || n >= 0x7fffffffull || stride > 0xffff0u) {
        SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
        if ((e = occupancy_grid(T, k_generic<false>, 256, L.total, &grid)) != cudaSuccess) return cudaGetErrorString(e);
        unsigned long long need = (n + 255) / 256; if ((unsigned long long)grid > need) grid = (int)need;
        k_generic<false><<<grid, 256, L.total, st>>>(T, (const uint8_t *)d_pkts, nullptr, stride, n, flags, d_verdicts);
        if (info) { info->grid = grid; info->block = 256; info->smem = L.total; }
        e = cudaGetLastError();
        return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
    }
    encode_tiled_fn enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled unavailable (driver too old?)";
    CUtensorMap map;
    cuuint64_t gdim[2] = { stride, n };
    cuuint64_t gstr[1] = { stride };
    cuuint32_t box[2] = { 64, 32 };                      // one warp's tile: 32 slots x 64 B
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(d_pkts), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     loader == LOADER_TMA_SW64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed";
#define FAST_ARGS T, map, d_pkts, stride, n, flags, d_verdicts, st, rare_q, rare_cnt, info
    const int minb = stages >> 4; stages &= 15;              // stages | minblocks << 4 (tuning knob CLK_STAGES / CLK_MINB)
    // every small table present is staged (and at least the endpoint table exists) -> shared-memory specialisation
    const bool all_staged = T.stage_ep_bytes && T.stage_soa_bytes && (!T.dns_b || T.stage_dns_bytes) && (!T.route_b || T.stage_route_bytes) &&
                            (!T.n_cidr || T.stage_pre_bytes) && (!T.ep6 || T.stage_ep6_bytes);
    if (loader != LOADER_TMA_SW64) return launch_fast<2, false, 1, false>(FAST_ARGS);
    if (all_staged) {
        if (minb >= 2) return launch_fast<1, true, 2, true>(FAST_ARGS);                 // 2 x 1024 threads per SM (<= 32 registers)
        return stages <= 1 ? launch_fast<1, true, 1, true>(FAST_ARGS) : launch_fast<2, true, 1, true>(FAST_ARGS);
    }
    if (stages <= 1) return launch_fast<1, true, 1, false>(FAST_ARGS);
    return launch_fast<2, true, 1, false>(FAST_ARGS);
#undef FAST_ARGS
}

static uint32_t varlen_slice() { uint32_t v = 12u << 20; if (const char *s = getenv("CLK_SLICE")) { long x = atol(s); if (x >= 1024) v = (uint32_t)x; } return v; }

template <int kType>
static cudaError_t launch_name_list(const DevTables &T, const void *d_pkts, uint4 *ents, uint32_t *cnt, uint32_t cap, uint32_t *d_verdicts, int sms, cudaStream_t st)
{
    static int nb = 0;                                                          // CTAs of 256 threads per SM (registers per thread follow): tuning knob CLK_NB
    if (!nb) { nb = 3; if (const char *s = getenv("CLK_NB")) { int v = atoi(s); if (v >= 2 && v <= 4) nb = v; } }
    const int grid = sms * nb;
    if (nb == 2) k_name_list<kType, 2><<<grid, 256, 0, st>>>(T, (const uint8_t *)d_pkts, ents, cnt, cap, d_verdicts);
    else if (nb == 4) k_name_list<kType, 4><<<grid, 256, 0, st>>>(T, (const uint8_t *)d_pkts, ents, cnt, cap, d_verdicts);
    else k_name_list<kType, 3><<<grid, 256, 0, st>>>(T, (const uint8_t *)d_pkts, ents, cnt, cap, d_verdicts);
    return cudaGetLastError();
}

// Work lists of one slice: 8 counters (padded to 256 bytes) + VW_LISTS lists of `cap` 16-byte entries; two such sets (slice
// parity), so the head kernel of slice j + 1 fills one while the parser kernels of slice j read the other.
static size_t varlen_set_bytes(uint32_t cap) { return 256 + (size_t)VW_LISTS * cap * 16; }
size_t varlen_scratch_bytes(unsigned long long n) { const uint32_t cap = (uint32_t)(n < varlen_slice() ? n : varlen_slice()); return 2 * varlen_set_bytes(cap) + 256; }

#define CU_TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cudaGetErrorString(_e); } while (0)

// The batch in slices; per slice: head kernel + generic list on `st`, then the parser kernels on three side streams (each
// runs one or two parsers; they are latency-bound and share the SMs with each other and with the NEXT slice's head kernel,
// which is why the head kernel is launched with one CTA per SM).  `st` joins every side stream before returning.
template <int kWarps, int kMinBlocks, int kPerSM>
static const char *launch_pipeline(const DevTables &T, const void *d_pkts, const uint32_t *d_offs16, unsigned long long n, uint32_t flags,
                                   uint32_t *d_verdicts, cudaStream_t st, const VarlenCtx *ctx, LaunchInfo *info)
{
    const uint32_t cap = (uint32_t)(n < varlen_slice() ? n : varlen_slice());
    if (!ctx || !ctx->scratch || ctx->scratch_bytes < 2 * varlen_set_bytes(cap)) return "varlen scratch too small";
    const int sms = sm_count() - reserved_sms(T);
    const SmemLayout L = smem_layout(T, 0, kWarps * VW_LISTS * VW_QCAP * 16, 0, MSLOTS, kWarps);
    CU_TRY(cudaFuncSetAttribute(k_varlen_head<kWarps, kMinBlocks>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const SmemLayout G = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    CU_TRY(cudaFuncSetAttribute(k_generic_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.total));
    const bool names = (flags & CLK_F_NAME_STAGE) != 0;
    uint32_t launches = 0, slice = 0;
    for (unsigned long long first = 0; first < n; first += cap, slice++) {
        const uint32_t ns = (uint32_t)(n - first < cap ? n - first : cap), p = slice & 1u;
        uint32_t *cnt = (uint32_t *)((uint8_t *)ctx->scratch + p * varlen_set_bytes(cap)); uint4 *ents = (uint4 *)((uint8_t *)cnt + 256);
        if (names && slice >= 2) for (int a = 0; a < 3; a++) CU_TRY(cudaStreamWaitEvent(st, ctx->ev_done[p][a], 0));     // the parsers of slice - 2 have left this set
        CU_TRY(cudaMemsetAsync(cnt, 0, 32, st));
        unsigned long long need = ((unsigned long long)(ns + 31) / 32 + kWarps - 1) / kWarps; const int g1 = (unsigned long long)(sms * kPerSM) > need ? (int)need : sms * kPerSM;
        k_varlen_head<kWarps, kMinBlocks><<<g1, kWarps * 32, L.total, st>>>(T, (const uint8_t *)d_pkts, d_offs16, (uint32_t)first, ns, flags, d_verdicts, ents, cnt, cap);
        CU_TRY(cudaGetLastError());
        if (info) { info->grid = g1; info->block = kWarps * 32; info->smem = L.total; }
        k_generic_list<<<sms, 256, G.total, st>>>(T, (const uint8_t *)d_pkts, ents, cnt, cap, flags, d_verdicts);
        CU_TRY(cudaGetLastError());
        launches += 2;
        if (!names) continue;
        CU_TRY(cudaEventRecord(ctx->ev_head[p], st));
        for (int a = 0; a < 3; a++) CU_TRY(cudaStreamWaitEvent(ctx->aux[a], ctx->ev_head[p], 0));
        CU_TRY(launch_name_list<(int)NM_DNSQ>(T, d_pkts, ents, cnt, cap, d_verdicts, sms, ctx->aux[0]));
        CU_TRY(launch_name_list<(int)NM_TLS>(T, d_pkts, ents, cnt, cap, d_verdicts, sms, ctx->aux[1]));
        if (T.n_http) { CU_TRY(launch_name_list<(int)NM_HTTP>(T, d_pkts, ents, cnt, cap, d_verdicts, sms, ctx->aux[1])); launches++; }
        CU_TRY(launch_name_list<(int)NM_DNSA>(T, d_pkts, ents, cnt, cap, d_verdicts, sms, ctx->aux[2]));
        launches += 3;
        for (int a = 0; a < 3; a++) CU_TRY(cudaEventRecord(ctx->ev_done[p][a], ctx->aux[a]));
    }
    if (names) {                                                                // join: the caller's stream sees everything done
        const uint32_t last = slice ? slice - 1 : 0;
        for (uint32_t q = 0; q < 2 && q <= last; q++) for (int a = 0; a < 3; a++) CU_TRY(cudaStreamWaitEvent(st, ctx->ev_done[(last - q) & 1u][a], 0));
    }
    if (info) info->launches = launches;
    return nullptr;
}

const char *launch_varlen(const DevTables &T, const void *d_pkts, const uint32_t *d_offs16, unsigned long long n, uint32_t flags,
                          uint32_t *d_verdicts, cudaStream_t st, const VarlenCtx *ctx, LaunchInfo *info)
{
    if (n == 0) return nullptr;
    int grid = 0; cudaError_t e;
    if (n < 0xfffffff0ull && !getenv("CLK_VARLEN_GENERIC")) {
        int cfgw = 0; if (const char *sw = getenv("CLK_VW")) cfgw = atoi(sw);
        switch (cfgw) {
        case 1:  return launch_pipeline<16, 2, 1>(T, d_pkts, d_offs16, n, flags, d_verdicts, st, ctx, info);   // 64 registers, one CTA per SM: leaves half an SM to the parsers
        case 2:  return launch_pipeline<8, 4, 2>(T, d_pkts, d_offs16, n, flags, d_verdicts, st, ctx, info);
        case 3:  return launch_pipeline<8, 4, 3>(T, d_pkts, d_offs16, n, flags, d_verdicts, st, ctx, info);
        default: return launch_pipeline<16, 2, 2>(T, d_pkts, d_offs16, n, flags, d_verdicts, st, ctx, info);
        }
    }
    SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    if ((e = occupancy_grid(T, k_generic<true>, 256, L.total, &grid)) != cudaSuccess) return cudaGetErrorString(e);
    unsigned long long need = (n + 255) / 256; if ((unsigned long long)grid > need) grid = (int)need;
    k_generic<true><<<grid, 256, L.total, st>>>(T, (const uint8_t *)d_pkts, d_offs16, 0, n, flags, d_verdicts);
    if (info) { info->grid = grid; info->block = 256; info->smem = L.total; }
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char *launch_events(const DevTables &T, const clk_event *d_ev, unsigned long long n, uint32_t flags, uint32_t *d_verdicts,
                          cudaStream_t st, LaunchInfo *info)
{
    if (n == 0) return nullptr;
    int grid = 0; cudaError_t e;
    SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    if ((e = occupancy_grid(T, k_events, 256, L.total, &grid)) != cudaSuccess) return cudaGetErrorString(e);
    unsigned long long need = (n + 255) / 256; if ((unsigned long long)grid > need) grid = (int)need;
    k_events<<<grid, 256, L.total, st>>>(T, d_ev, n, flags, d_verdicts);
    if (info) { info->grid = grid; info->block = 256; info->smem = L.total; }
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char *launch_hdr16(const DevTables &T, const void *d_recs, const void *d_aux, unsigned long long n, uint32_t flags, uint32_t *d_verdicts,
                         cudaStream_t st, LaunchInfo *info)
{
    if (n == 0) return nullptr;
    int grid = 0; cudaError_t e;
    SmemLayout L = smem_layout(T, 0, 0, 0, MSLOTS, 8);
    if ((e = occupancy_grid(T, k_hdr16, 256, L.total, &grid)) != cudaSuccess) return cudaGetErrorString(e);
    unsigned long long need = (n + 255) / 256; if ((unsigned long long)grid > need) grid = (int)need;
    k_hdr16<<<grid, 256, L.total, st>>>(T, (const uint4 *)d_recs, (const uint8_t *)d_aux, n, flags, d_verdicts);
    if (info) { info->grid = grid; info->block = 256; info->smem = L.total; info->launches = 1; }
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char *launch_dns_extract(const DevTables &T, const void *d_pkts, const uint32_t *d_offs16, unsigned long long n,
                               const uint32_t *d_verdicts, unsigned long long *d_block_sums, unsigned long long *d_pos,
                               unsigned long long *d_total, clk_dns_update_rec *d_out, unsigned long long cap, cudaStream_t st)
{
    if (n == 0) return nullptr;
    uint32_t nb = (uint32_t)((n + 4095) / 4096);
    k_count_block<<<nb, 256, 0, st>>>(d_verdicts, n, d_block_sums);
    k_scan_blocks<<<1, 32, 0, st>>>(d_block_sums, nb, d_total);
    k_positions<<<nb, 256, 0, st>>>(d_verdicts, n, d_block_sums, d_pos);
    int grid = sm_count() * 4; unsigned long long need = (n + 255) / 256; if ((unsigned long long)grid > need) grid = (int)need;
    k_dns_extract<<<grid, 256, 0, st>>>(T, (const uint8_t *)d_pkts, d_offs16, n, d_verdicts, d_pos, d_out, cap);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char *launch_counters_export(const DevTables &T, clk_metric *d_block, uint32_t cap_rows, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_block, 0, sizeof(clk_metric), st);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    k_counters_export<<<side_grid(T, sm_count() * 8), 256, 0, st>>>(T, d_block, cap_rows);
    e = cudaGetLastError(); return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
const char *launch_counters_totals(const DevTables &T, clk_total *d_block, uint32_t rows, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_block, 0, (size_t)rows * sizeof(clk_total), st);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    const uint32_t nr = T.n_crows < rows ? T.n_crows : rows;
    const size_t smem = nr * 24u <= 48u * 1024u ? nr * 24u : 0;
    k_counters_totals<<<side_grid(T, sm_count() * 4), 256, smem, st>>>(T, d_block, rows);
    e = cudaGetLastError(); return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
const char *launch_counters_clear(const DevTables &T, cudaStream_t st)
{
    k_counters_clear<<<side_grid(T, sm_count() * 8), 256, 0, st>>>(T);
    cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
const char *launch_counters_import(const DevTables &T, const clk_metric *d_blocks, uint32_t n_blocks, uint32_t block_rows, cudaStream_t st)
{
    unsigned long long blocks = ((unsigned long long)n_blocks * block_rows + 255) / 256;   // one row per thread: the adds are L2-latency bound
    if (blocks > (unsigned long long)sm_count() * 64) blocks = (unsigned long long)sm_count() * 64;
    if (blocks < 1) blocks = 1;
    k_counters_import<<<side_grid(T, (int)blocks), 256, 0, st>>>(T, d_blocks, n_blocks, block_rows);
    cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace clk
