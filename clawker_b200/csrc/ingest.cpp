// ingest.cpp — the capture front end either side of the classifier (no reference counterpart: the reference's
// hooks see sockets, never frames).
//
//   clk_afpacket_*    live ingest: an AF_PACKET / TPACKET_V3 receive ring on one interface, drained block by block into the
//                     varlen layout of clk_classify_varlen (frames padded to >= 64 B and to 16 B, u32 offsets in 16-byte units).
//   clk_tcp_coalesce  TCP reassembly of the one thing the name stage needs it for: a TLS ClientHello or a plaintext HTTP request
//                     head that does not fit its first segment (large / post-quantum key shares, long header blocks).  Envoy's
//                     tls_inspector / HTTP codec buffer the connection until they can decide; a per-packet classifier says N/A
//                     ("needs reassembly") for such a first segment.  This pass appends the in-order continuation segments of the
//                     same flow to the first segment's frame (so the classifier sees the whole record) and empties the
//                     continuation segments (headers only: no payload, name verdict N/A), keeping one output frame per input
//                     frame, so verdict arrays stay aligned with the capture.  IPv4 / IHL-any TCP only (the name stage's own scope).
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <arpa/inet.h>
#include <linux/if_packet.h>
#include <net/ethernet.h>
#include <net/if.h>
#include <poll.h>
#include <sys/ioctl.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <unistd.h>

#include "../../include/clawker_gpu.h"
#include "hdr16_pack.h"

namespace {
thread_local std::string tl_ingest_error;
int ifail(int code, const std::string &m) { tl_ingest_error = m; return code; }

inline uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
}  // namespace

struct clk_ring {
    int fd = -1; uint8_t *map = nullptr; size_t map_len = 0;
    uint32_t block_size = 0, block_nr = 0, cur = 0;
    uint32_t pkt_in_block = 0; uint8_t *pkt_ptr = nullptr;          // resume point inside a block the caller's buffers could not take whole
    uint64_t frames = 0, dropped = 0;
};

#define ING_GUARD catch (const std::bad_alloc &) { return ifail(CLK_ENOMEM, "out of memory"); } \
    catch (const std::exception &x) { return ifail(CLK_EIO, std::string("internal error: ") + x.what()); } catch (...) { return ifail(CLK_EIO, "internal error"); }

extern "C" {

const char *clk_ingest_last_error(void) { return tl_ingest_error.c_str(); }

int clk_afpacket_open(const char *ifname, uint32_t ring_mib, clk_ring **out)
try {
    if (!ifname || !out) return ifail(CLK_EINVAL, "NULL");
    *out = nullptr;
    const unsigned ifi = if_nametoindex(ifname);
    if (!ifi) return ifail(CLK_ENOENT, std::string("no such interface: ") + ifname);
    int fd = socket(AF_PACKET, SOCK_RAW, htons(ETH_P_ALL));
    if (fd < 0) return ifail(CLK_EIO, std::string("socket(AF_PACKET): ") + strerror(errno) + " (needs CAP_NET_RAW)");
    int ver = TPACKET_V3;
    struct tpacket_req3 req; memset(&req, 0, sizeof req);
    req.tp_block_size = 1u << 20; req.tp_frame_size = 2048;
    req.tp_block_nr = ring_mib ? ring_mib : 64; req.tp_frame_nr = req.tp_block_nr * (req.tp_block_size / req.tp_frame_size);
    req.tp_retire_blk_tov = 10;                                       // ms: a partly filled block is handed over after this long
    struct sockaddr_ll sll; memset(&sll, 0, sizeof sll);
    sll.sll_family = AF_PACKET; sll.sll_protocol = htons(ETH_P_ALL); sll.sll_ifindex = (int)ifi;
    if (setsockopt(fd, SOL_PACKET, PACKET_VERSION, &ver, sizeof ver) < 0 || setsockopt(fd, SOL_PACKET, PACKET_RX_RING, &req, sizeof req) < 0) {
        const std::string m = std::string("PACKET_RX_RING: ") + strerror(errno); close(fd); return ifail(CLK_EIO, m);
    }
    const size_t len = (size_t)req.tp_block_size * req.tp_block_nr;
    void *m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_LOCKED, fd, 0);
    if (m == MAP_FAILED) m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { const std::string e = std::string("mmap ring: ") + strerror(errno); close(fd); return ifail(CLK_ENOMEM, e); }
    {   // a loopback interface hands every packet to a tap twice (outgoing + incoming): keep the incoming copy
        struct ifreq ifr; memset(&ifr, 0, sizeof ifr); strncpy(ifr.ifr_name, ifname, IFNAMSIZ - 1);
        if (ioctl(fd, SIOCGIFFLAGS, &ifr) == 0 && (ifr.ifr_flags & IFF_LOOPBACK)) {
#ifdef PACKET_IGNORE_OUTGOING
            int one = 1; setsockopt(fd, SOL_PACKET, PACKET_IGNORE_OUTGOING, &one, sizeof one);
#endif
        }
    }
    if (bind(fd, (struct sockaddr *)&sll, sizeof sll) < 0) { const std::string e = std::string("bind: ") + strerror(errno); munmap(m, len); close(fd); return ifail(CLK_EIO, e); }
    clk_ring *r = new clk_ring();
    r->fd = fd; r->map = (uint8_t *)m; r->map_len = len; r->block_size = req.tp_block_size; r->block_nr = req.tp_block_nr;
    *out = r;
    return CLK_OK;
} ING_GUARD

int clk_afpacket_close(clk_ring *r)
{
    if (!r) return CLK_OK;
    if (r->map) munmap(r->map, r->map_len);
    if (r->fd >= 0) close(r->fd);
    delete r;
    return CLK_OK;
}

/* Drains the blocks the kernel has handed over into (blob, offs16) until a capacity runs out or no block is ready (after waiting
 * up to timeout_ms for the first one).  *n_frames frames were written; offs16[0 .. n_frames] is set. */
int clk_afpacket_read(clk_ring *r, uint8_t *blob, uint64_t cap_bytes, uint32_t *offs16, uint64_t cap_frames, int timeout_ms, uint64_t *n_frames)
try {
    if (!r || !blob || !offs16 || !n_frames) return ifail(CLK_EINVAL, "NULL");
    uint64_t n = 0, used = 0; bool waited = false;
    while (n < cap_frames) {
        struct tpacket_block_desc *bd = (struct tpacket_block_desc *)(r->map + (size_t)r->cur * r->block_size);
        if (!(bd->hdr.bh1.block_status & TP_STATUS_USER)) {
            if (n || waited) break;
            struct pollfd pfd = { r->fd, POLLIN | POLLERR, 0 };
            waited = true;
            if (poll(&pfd, 1, timeout_ms) <= 0) break;
            continue;
        }
        if (!r->pkt_ptr) { r->pkt_in_block = 0; r->pkt_ptr = (uint8_t *)bd + bd->hdr.bh1.offset_to_first_pkt; }
        const uint32_t num = bd->hdr.bh1.num_pkts;
        bool full = false;
        while (r->pkt_in_block < num) {
            const struct tpacket3_hdr *ph = (const struct tpacket3_hdr *)r->pkt_ptr;
            const uint32_t caplen = ph->tp_snaplen;
            const uint64_t slot = caplen < 64 ? 64 : ((uint64_t)caplen + 15) & ~15ull;
            if (n >= cap_frames || used + slot > cap_bytes || ((used + slot) / 16) >> 32) { full = true; break; }
            memcpy(blob + used, r->pkt_ptr + ph->tp_mac, caplen);
            memset(blob + used + caplen, 0, slot - caplen);
            offs16[n] = (uint32_t)(used / 16);
            used += slot; n++;
            r->pkt_in_block++;
            r->pkt_ptr += ph->tp_next_offset ? ph->tp_next_offset : 0;
            if (!ph->tp_next_offset) { r->pkt_in_block = num; }
        }
        if (full) {
            if (n == 0) return ifail(CLK_E2BIG, "the next frame does not fit the buffers");      // (a caller looping on n == 0 would spin)
            break;
        }
        bd->hdr.bh1.block_status = TP_STATUS_KERNEL;                    // the block goes back to the kernel
        __sync_synchronize();
        r->pkt_ptr = nullptr; r->cur = (r->cur + 1) % r->block_nr;
    }
    offs16[n] = (uint32_t)(used / 16);
    *n_frames = n; r->frames += n;
    return CLK_OK;
} ING_GUARD

/* The same drain, but straight into compact ingest records (clk_hdr16, + 64-byte aux slots for the frames a record cannot express):
 * what clk_classify_hdr16 takes, 16 bytes per frame over the host link.  Stops when recs or aux is full (the block is resumed by the
 * next call).  No payload leaves the ring on this path: no name stage. */
int clk_afpacket_read_hdr16(clk_ring *r, clk_hdr16 *recs, uint64_t cap_frames, uint8_t *aux, uint64_t aux_cap, int timeout_ms,
                            uint64_t *n_frames, uint64_t *n_aux)
try {
    if (!r || !recs || !n_frames || !n_aux || (aux_cap && !aux)) return ifail(CLK_EINVAL, "NULL");
    uint64_t n = 0, na = 0; bool waited = false;
    while (n < cap_frames) {
        struct tpacket_block_desc *bd = (struct tpacket_block_desc *)(r->map + (size_t)r->cur * r->block_size);
        if (!(bd->hdr.bh1.block_status & TP_STATUS_USER)) {
            if (n || waited) break;
            struct pollfd pfd = { r->fd, POLLIN | POLLERR, 0 };
            waited = true;
            if (poll(&pfd, 1, timeout_ms) <= 0) break;
            continue;
        }
        if (!r->pkt_ptr) { r->pkt_in_block = 0; r->pkt_ptr = (uint8_t *)bd + bd->hdr.bh1.offset_to_first_pkt; }
        const uint32_t num = bd->hdr.bh1.num_pkts;
        bool full = false;
        while (r->pkt_in_block < num) {
            const struct tpacket3_hdr *ph = (const struct tpacket3_hdr *)r->pkt_ptr;
            if (n >= cap_frames) { full = true; break; }
            uint8_t head[64];
            const uint32_t take = ph->tp_snaplen < 64 ? ph->tp_snaplen : 64;
            memcpy(head, r->pkt_ptr + ph->tp_mac, take); if (take < 64) memset(head + take, 0, 64 - take);
            clk_hdr16 rec;
            if (clk_hdr16_from_frame(head, rec)) {
                if (na >= aux_cap || na > 0xffffffffull) { full = true; break; }
                memcpy(aux + na * 64, head, 64);
                rec.kind = CLK_HDR_AUX; rec.src = (uint32_t)na++;
            }
            recs[n++] = rec;
            r->pkt_in_block++;
            if (!ph->tp_next_offset) r->pkt_in_block = num; else r->pkt_ptr += ph->tp_next_offset;
        }
        if (full) {
            if (n == 0) return ifail(CLK_E2BIG, "no room for a record (or its aux slot)");
            break;
        }
        bd->hdr.bh1.block_status = TP_STATUS_KERNEL;
        __sync_synchronize();
        r->pkt_ptr = nullptr; r->cur = (r->cur + 1) % r->block_nr;
    }
    *n_frames = n; *n_aux = na; r->frames += n;
    return CLK_OK;
} ING_GUARD

/* See the header of this file.  in: (blob, offs16, n) in the varlen layout.  out_blob must hold in-bytes + the bytes the merges
 * add (<= in-bytes again is always enough); out_offs16 has n + 1 entries.  *n_merged = continuation segments folded into a first
 * segment. */
int clk_tcp_coalesce(const uint8_t *blob, const uint32_t *offs16, uint64_t n, uint8_t *out_blob, uint64_t out_cap_bytes,
                     uint32_t *out_offs16, uint64_t *n_merged)
try {
    if ((n && (!blob || !offs16)) || !out_blob || !out_offs16) return ifail(CLK_EINVAL, "NULL");
    struct Seg { uint32_t l4off, payoff, paylen, seq; bool tcp; };
    struct Pending { uint64_t first; uint32_t next_seq, have, need; uint32_t age; bool http; };
    auto parse = [&](uint64_t i, Seg &s) -> const uint8_t * {
        const uint8_t *p = blob + (uint64_t)offs16[i] * 16; const uint32_t avail = (offs16[i + 1] - offs16[i]) * 16;
        s = Seg{0, 0, 0, 0, false};
        if (avail < 54 || p[12] != 0x08 || p[13] != 0x00 || (p[14] >> 4) != 4) return p;
        const uint32_t ihl = p[14] & 15u; if (ihl < 5 || p[23] != 6 || (be16(p + 20) & 0x1fffu)) return p;
        s.l4off = 14 + 4 * ihl; if (s.l4off + 20 > avail) return p;
        const uint32_t doff = (p[s.l4off + 12] >> 4) * 4; if (doff < 20) return p;
        uint32_t end = 14 + be16(p + 16); if (end > avail) end = avail;
        s.payoff = s.l4off + doff; if (s.payoff > end) return p;
        s.paylen = end - s.payoff; s.seq = be32(p + s.l4off + 4); s.tcp = true;
        return p;
    };
    struct KeyHash { size_t operator()(const std::pair<uint64_t, uint32_t> &k) const { return (size_t)(k.first * 0x9E3779B97F4A7C15ull ^ k.second * 0x85EBCA77u); } };
    std::unordered_map<std::pair<uint64_t, uint32_t>, Pending, KeyHash> flows;
    std::vector<std::vector<uint64_t>> tails;                            // per merge: the continuation packets in order
    std::vector<int64_t> role(n, -1);                                    // -1 copy as is; >= 0: first segment of merge #role; -2: continuation (emptied)
    uint64_t merged = 0;
    for (uint64_t i = 0; i < n; i++) {
        Seg s; const uint8_t *p = parse(i, s);
        if (!s.tcp) continue;
        const std::pair<uint64_t, uint32_t> key{ ((uint64_t)be32(p + 26) << 32) | be32(p + 30), (be16(p + s.l4off) << 16) | be16(p + s.l4off + 2) };
        const uint8_t flags = p[s.l4off + 13];
        auto it = flows.find(key);
        if (it != flows.end()) {
            Pending &pd = it->second;
            if ((flags & 0x05) || i - pd.first > 4096) { flows.erase(it); }                    // FIN / RST, or the flow went quiet: give up (the first segment stays N/A)
            else if (s.paylen && s.seq == pd.next_seq) {
                role[i] = -2; tails[(size_t)role[pd.first]].push_back(i); merged++;
                pd.next_seq += s.paylen; pd.have += s.paylen;
                bool done = pd.have >= pd.need || pd.have > (1u << 16);
                if (pd.http && !done) {                                                          // a request head ends at the first empty line
                    const uint8_t *q = p + s.payoff;
                    for (uint32_t k = 0; k + 3 < s.paylen && !done; k++) done = q[k] == '\r' && q[k + 1] == '\n' && q[k + 2] == '\r' && q[k + 3] == '\n';
                }
                if (done) flows.erase(it);
                continue;
            } else if (s.paylen) continue;                                                       // out of order / retransmission: not merged
            else continue;
        }
        if (!s.paylen) continue;
        const uint8_t *q = p + s.payoff;
        bool tls = s.paylen >= 5 && q[0] == 0x16 && q[1] == 0x03 && 5 + be16(q + 3) > s.paylen;
        bool http = false;
        if (!tls && q[0] >= 'A' && q[0] <= 'Z') {                                               // a request line without the end of the header block
            uint32_t k = 0; while (k < s.paylen && k < 16 && q[k] >= 'A' && q[k] <= 'Z') k++;
            if (k >= 3 && k < s.paylen && q[k] == ' ') {
                http = true;
                for (uint32_t j = 0; j + 3 < s.paylen && http; j++) if (q[j] == '\r' && q[j + 1] == '\n' && q[j + 2] == '\r' && q[j + 3] == '\n') http = false;
            }
        }
        if (!tls && !http) continue;
        Pending pd; pd.first = i; pd.next_seq = s.seq + s.paylen; pd.have = s.paylen; pd.need = tls ? 5 + be16(q + 3) : 0xffffffffu; pd.age = 0; pd.http = http;
        role[i] = (int64_t)tails.size(); tails.emplace_back();
        flows[key] = pd;
    }
    for (uint64_t i = 0; i < n; i++) if (role[i] >= 0 && tails[(size_t)role[i]].empty()) role[i] = -1;     // nothing followed: the frame stays as it is
    uint64_t used = 0;
    for (uint64_t i = 0; i < n; i++) {
        Seg s; const uint8_t *p = parse(i, s); const uint32_t avail = (offs16[i + 1] - offs16[i]) * 16;
        uint64_t len = avail;
        if (role[i] == -2) len = s.payoff;                                                       // headers only
        else if (role[i] >= 0) { len = (uint64_t)s.payoff + s.paylen; for (uint64_t j : tails[(size_t)role[i]]) { Seg t; parse(j, t); len += t.paylen; } }
        const uint64_t slot = len < 64 ? 64 : (len + 15) & ~15ull;
        if (used + slot > out_cap_bytes || ((used + slot) / 16) >> 32) return ifail(CLK_E2BIG, "output blob too small");
        uint8_t *o = out_blob + used;
        if (role[i] == -1) { memcpy(o, p, avail); len = avail; }
        else {
            memcpy(o, p, s.payoff);
            uint64_t w = s.payoff;
            if (role[i] >= 0) {
                memcpy(o + w, p + s.payoff, s.paylen); w += s.paylen;
                for (uint64_t j : tails[(size_t)role[i]]) { Seg t; const uint8_t *pj = parse(j, t); memcpy(o + w, pj + t.payoff, t.paylen); w += t.paylen; }
            }
            const uint64_t ip_total = w - 14 > 0xffff ? 0xffff : w - 14;                        // a merged frame may exceed 64 KiB: the IP length saturates, the classifier clamps to the captured bytes
            o[16] = (uint8_t)(ip_total >> 8); o[17] = (uint8_t)ip_total;
            len = w;
        }
        if (slot > len) memset(o + len, 0, slot - len);
        out_offs16[i] = (uint32_t)(used / 16);
        used += slot;
    }
    out_offs16[n] = (uint32_t)(used / 16);
    if (n_merged) *n_merged = merged;
    return CLK_OK;
} ING_GUARD

}  // extern "C"
