// dev_tables.h — immutable device snapshot of the firewall tables (host builds, kernels read).
//
// Layouts mirror the reference's BPF maps (internal/ebpf/bpf/common.h:80-177) flattened into
// open-addressing tables whose slots are single 8/16-byte words so that one 64-/128-bit load
// answers a lookup.  Everything here is POD and shared verbatim between engine.cpp (host) and
// classify.cu (device); hash functions are __host__ __device__ so both sides agree.
#pragma once
#include <stdint.h>
#include "../../include/clawker_gpu.h"

#if defined(__CUDACC__)
#define CLK_HD __host__ __device__ __forceinline__
#else
#define CLK_HD inline
#endif

namespace clk {

// ---- hashing --------------------------------------------------------------------------------
CLK_HD uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
#if !defined(__CUDACC__)
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
#endif
CLK_HD uint32_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}

// Name hash: two independent 32-bit polynomial hashes over the lower-cased bytes, low power first:
//   h(s) = sum s[j] * B^j  (mod 2^32),  B in {NAME_B1, NAME_B2}.
// A right-to-left Horner pass yields h of every suffix as it goes; a left-to-right pass yields them through
// h(s[k:]) = (F - P_k) * B^-k with F the whole-string sum and P_k the prefix sum (B is odd, so invertible).
// A hit is always confirmed byte-for-byte against the stored name, so the hash only has to be a good filter
// (collisions cost time, never correctness).
static const uint32_t NAME_B1 = 0x9E3779B1u, NAME_B1INV = 0x0E8B2F51u;      // B1 * B1INV == 1 (mod 2^32)
static const uint32_t NAME_B2 = 0x85EBCA77u, NAME_B2INV = 0xB6C92F47u;
CLK_HD uint64_t name_key(uint32_t h1, uint32_t h2, uint32_t len)
{
    uint64_t k = (((uint64_t)h2 << 32) | h1) ^ ((uint64_t)len * 0xD6E8FEB86659FD93ull);
    return k ? k : 1;                                              // 0 is the empty-slot marker
}

// Name tables are open addressing over 16-byte NameSlots with linear probing, HOME SLOTS EVEN and load <= 1/4: the
// pair {home, home + 1} is one aligned 32-byte word, and a lookup that finds neither its key nor a full pair there is
// decided by that one load (an empty slot ends every probe sequence; nothing is ever deleted — tables are rebuilt).
CLK_HD uint32_t name_home(uint64_t key, uint32_t mask)
{
    uint32_t x = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA77u);
    x ^= x >> 15;
    return x & mask & ~1u;
}

// ---- rows ------------------------------------------------------------------------------------
#define CLK_ROW_ENFORCED 1u     // cgroup present in container_map (common.h:273)
#define CLK_ROW_BYPASS   2u     // bypass_map[cg] == 1             (common.h:266-267)

// One endpoint (packet source address -> socket identity + its container's config), 64 B.
struct EpRow {
    uint32_t ip;                // IPv4 key (0 for IPv6 endpoints)
    uint32_t uid;
    uint64_t cgroup_id;
    uint32_t so_mark;
    uint32_t flags;             // CLK_ROW_*
    clk_container_config cfg;   // 28 B copy (valid when ENFORCED)
    uint32_t crow;              // index of the cgroup's ContRow (metric key); 0xffffffff none
    uint32_t ip6[2];            // low 8 bytes of an IPv6 key are in ip6x (see Ep6Key)
};
static_assert(sizeof(EpRow) == 64, "EpRow must be 64 B");

// One cgroup (events API lookup + metric key -> cgroup_id), 48 B.
struct ContRow {
    uint64_t cgroup_id;
    uint32_t flags;
    clk_container_config cfg;
    uint32_t _pad[2];
};
static_assert(sizeof(ContRow) == 48, "ContRow must be 48 B");

struct Ep6Key { uint32_t a[4]; uint32_t row; uint32_t _pad[3]; };   // 32 B; row = index into EpRow
static_assert(sizeof(Ep6Key) == 32, "Ep6Key");

// Name table slot (zones and SNI server_names), 16 B: key==0 empty.
struct NameSlot {
    uint64_t key;               // name_key(hash, len)
    uint32_t val;               // zones: DomainHash(zone) (types.go:121-125); SNI: entry index
    uint32_t str_off;           // length << 24 | (offset into the names blob) / 16: lower-case bytes, 16-byte aligned, zero padded
};

enum { HTTP_K_EXACT = 0, HTTP_K_SUFFIX = 1, HTTP_K_PREFIX = 2 };

// ---- dns_cache / route_map: 2-choice bucketed tables -------------------------------------------
// A bucket is one 16-byte word holding two 8-byte slots, so a lookup is one 128-bit load and a
// miss is decided without a probe loop: keys go to bucket b1 while it has room, else to b2.
//   dns slot  : lo32 = ip (key), hi32 = domain_hash.                       empty = ~0ull
//   route slot: [63:32] domain_hash, [31:16] dst_port, [15:0] envoy_port.  empty = ~0ull
static const uint64_t SLOT_EMPTY = ~0ull;
// Table hashes are a single multiply whose TOP bits index the table (index = hash >> shift, with
// shift = 32 - log2(size)): the tables are rebuilt (and grown) by the host until every key fits, so
// hash quality only affects table size, never correctness.
CLK_HD uint32_t hash_a(uint32_t x) { return x * 0x9E3779B1u; }
CLK_HD uint32_t hash_b(uint32_t x) { return (x ^ 0x68E31DA4u) * 0x85EBCA77u; }
CLK_HD uint32_t route_fold(uint32_t hash, uint32_t port) { return hash ^ (port * 0x01000193u); }
CLK_HD uint32_t shift_for(uint32_t size_pow2) { uint32_t k = 0; while ((1u << k) < size_pow2) k++; return 32u - k; }

// Endpoint rows again as structure-of-arrays (one u32 array per field, `soa_stride` apart) so that a
// warp reading one field of 32 different rows from shared memory hits 32 different banks.
enum { SOA_UID = 0, SOA_MARK, SOA_FLAGS, SOA_ENVOY, SOA_COREDNS, SOA_GATEWAY, SOA_NET, SOA_MASK, SOA_HPIP, SOA_PORTS /* hp_port | egress << 16 */, SOA_CROW, SOA_FIELDS = 12 };

// ---- metrics (per-GPU metrics_map) -----------------------------------------------------------
// Packed key: [31:0] domain_hash, [47:32] dst_port, [61:48] crow, [63:62] action+1 (never 0).
CLK_HD uint64_t metric_pack(uint32_t hash, uint32_t port, uint32_t crow, uint32_t action)
{
    return (uint64_t)hash | ((uint64_t)(port & 0xffffu) << 32) | ((uint64_t)(crow & 0x3fffu) << 48) |
           ((uint64_t)(action + 1) << 62);
}
#define CLK_MAX_CROWS 16384u

// ---- the snapshot handed to kernels by value -------------------------------------------------
struct DevTables {
    const EpRow    *ep_rows;     uint32_t n_ep_rows;
    const uint2    *ep_slots;    uint32_t ep_mask, ep_shift;   // slot: {ip, EpRow index + 1} (y == 0 empty), index hash_a(ip) >> ep_shift, load <= 1/4
    const uint4    *ep_rowq;                             // 3 x uint4 per EpRow index: {uid,mark,flags,crow} {envoy,coredns,gateway,hp_ip} {net,mask,hp_port|egress<<16,0}
    const Ep6Key   *ep6;         uint32_t n_ep6, ep6_mask, ep6_shift;   // open addressing, 32 B slots {addr[4], row}; row == 0xffffffff empty;
                                                                         // index hash_a(a0^a1^a2^a3) >> ep6_shift
    const ContRow  *crows;       uint32_t n_crows;
    const uint32_t *cg_slots;    uint32_t cg_mask;       // slot: ContRow index + 1, keyed by mix64(cgroup_id)
    const uint4    *dns_b;       uint32_t dns_shift;     // buckets; null => empty table; index hash_{a,b}(ip) >> dns_shift
    const uint4    *route_b;     uint32_t route_shift;
    uint32_t dns_ffff_present;   uint32_t dns_ffff_hash; // the one key/value pair that equals SLOT_EMPTY
    // CIDR allow set (A12): 2 bits per /24 block: bit0 = covered by a prefix <= /24,
    // bit1 = block has longer prefixes -> cidr_l2 (hash by block -> 256-bit host bitmap)
    const uint32_t *cidr_l1;     uint32_t n_cidr;        // n_cidr == 0 => stage skipped
    const uint32_t *cidr_pre;                            // 64 Kbit prefilter: bit (addr >> 16) set iff some rule intersects that /16
    const uint32_t *cidr_l2_keys; const uint32_t *cidr_l2_bits; uint32_t cidr_l2_mask, cidr_l2_shift;   // key: block+1; bits: 8 words / slot
    const NameSlot *zone_slots;  uint32_t zone_mask;     // mask = slots - 1; home = name_home(key, mask)
    const NameSlot *sni_slots;   uint32_t sni_mask;
    uint32_t zone_max_labels;                            // most labels of any zone: longer suffixes of a query name cannot be zones
    uint32_t sni_max_dots;                               // most '.' in any wildcard server_name ".d": longer dot-suffixes cannot match
    // F3: plaintext HTTP virtual hosts.  Table strings: exact "d" | suffix ".d" (of "*.d") | prefix "d:" (of "d:*");
    // val = virtual-host index | kind << 28.  http_vh[v] = names-blob offset of the route list
    // [n u8] n x { allow u8, len u8, prefix bytes }.
    const NameSlot *http_slots;  uint32_t http_mask;
    const uint32_t *http_vh;     uint32_t n_http;
    const uint8_t  *names;                               // blob for string confirmation
    uint32_t coredns_ips[8];     uint32_t n_coredns;     // distinct cfg.coredns_ip values (DNS-answer stage)
    // device metrics table
    unsigned long long *m_keys;  unsigned long long *m_counts; uint32_t m_mask;
    unsigned long long *m_dropped;                       // [0] increments dropped, [1] keys present
    uint32_t m_max;                                      // max_entries of metrics_map (common.h:173): inserts beyond it are dropped
    // sizes for optional shared-memory staging (bytes; 0 = do not stage)
    uint32_t stage_ep_bytes, stage_soa_bytes, stage_dns_bytes, stage_route_bytes, stage_pre_bytes, stage_ep6_bytes;
    uint32_t reserved_sms;                               // SMs the persistent classify kernels leave to the side-stream collective (clk_cfg)
};

}  // namespace clk
