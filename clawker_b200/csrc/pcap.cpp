// pcap.cpp — capture-file ingest for the batch classifier: classic pcap -> the varlen layout
// (frames padded to >= 64 B and to 16 B, u32 offsets in 16 B units) in pinned host memory, double-buffered so
// that reading and packing chunk k+1 overlaps the H2D / kernel / D2H pipeline of chunk k (clk_classify_varlen).
// The reference has no capture path (its hooks see sockets, not frames); the on-wire format is the one
// tcpdump / AF_PACKET tooling writes: pcap magic a1b2c3d4 (microseconds) or a1b23c4d (nanoseconds), either
// byte order, LINKTYPE_ETHERNET.
#include <cstdio>
#include <cstring>
#include <future>
#include <string>
#include <vector>

#include "../../include/clawker_gpu.h"

namespace {

thread_local std::string tl_pcap_error;

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

struct PcapFile {
    FILE *f = nullptr; bool swapped = false; uint32_t snaplen = 0; uint64_t size = 0;
    ~PcapFile() { if (f) fclose(f); }
    int open(const char *path, uint64_t cursor)
    {
        f = fopen(path, "rb");
        if (!f) { tl_pcap_error = std::string("open ") + path + ": " + strerror(errno); return errno == ENOENT ? CLK_ENOENT : CLK_EIO; }
        fseeko(f, 0, SEEK_END); size = (uint64_t)ftello(f); fseeko(f, 0, SEEK_SET);
        uint32_t h[6];
        if (fread(h, 1, 24, f) != 24) { tl_pcap_error = std::string(path) + ": shorter than a pcap header"; return CLK_EINVAL; }
        if (h[0] == 0xA1B2C3D4u || h[0] == 0xA1B23C4Du) swapped = false;
        else if (h[0] == 0xD4C3B2A1u || h[0] == 0x4D3CB2A1u) swapped = true;
        else { tl_pcap_error = std::string(path) + ": unsupported pcap magic " + std::to_string(h[0]); return CLK_EINVAL; }
        snaplen = swapped ? bswap32(h[4]) : h[4];
        const uint32_t link = swapped ? bswap32(h[5]) : h[5];
        if ((link & 0x0fffffffu) != 1u) { tl_pcap_error = std::string(path) + ": link type " + std::to_string(link) + " is not Ethernet"; return CLK_EINVAL; }
        if (cursor) { if (cursor < 24 || cursor > size) { tl_pcap_error = "cursor outside the file"; return CLK_EINVAL; } fseeko(f, (off_t)cursor, SEEK_SET); }
        return CLK_OK;
    }
};

// Packs records until a capacity runs out or the file ends; a truncated last record ends the file.
int pack(PcapFile &pf, uint8_t *blob, uint64_t cap_bytes, uint32_t *offs16, uint64_t cap_frames, uint64_t base16, uint64_t *n_frames, uint64_t *n_bytes, bool *eof)
{
    uint64_t n = 0, used = 0; *eof = false;
    while (n < cap_frames) {
        const off_t rec = ftello(pf.f);
        uint32_t rh[4];
        if (fread(rh, 1, 16, pf.f) != 16) { *eof = true; fseeko(pf.f, rec, SEEK_SET); break; }
        const uint32_t caplen = pf.swapped ? bswap32(rh[2]) : rh[2];
        if ((uint64_t)rec + 16 + caplen > pf.size || caplen > 0x4000000u) { *eof = true; fseeko(pf.f, rec, SEEK_SET); break; }   // truncated dump
        const uint64_t slot = caplen < 64 ? 64 : ((uint64_t)caplen + 15) & ~15ull;
        if (used + slot > cap_bytes || ((base16 + (used + slot) / 16) >> 32)) { fseeko(pf.f, rec, SEEK_SET); break; }
        if (fread(blob + used, 1, caplen, pf.f) != caplen) { *eof = true; fseeko(pf.f, rec, SEEK_SET); break; }
        memset(blob + used + caplen, 0, slot - caplen);
        offs16[n] = (uint32_t)(base16 + used / 16);
        used += slot; n++;
    }
    offs16[n] = (uint32_t)(base16 + used / 16);
    *n_frames = n; *n_bytes = used;
    return CLK_OK;
}

}  // namespace

// Exception barrier (see engine.cpp): std::async / std::string can throw; nothing may unwind into the caller.
#define CLK_GUARD_INT(e) catch (const std::bad_alloc &) { tl_pcap_error = "out of memory"; return CLK_ENOMEM; } \
    catch (const std::exception &x) { tl_pcap_error = std::string("internal error: ") + x.what(); return CLK_EIO; } catch (...) { tl_pcap_error = "internal error"; return CLK_EIO; }

extern "C" {

const char *clk_pcap_last_error(void) { return tl_pcap_error.c_str(); }

int clk_pcap_pack(const char *path, uint64_t *cursor, uint8_t *blob, uint64_t cap_bytes, uint32_t *offs16, uint64_t cap_frames, uint64_t *n_frames)
try {
    if (!path || !cursor || !blob || !offs16 || !n_frames) { tl_pcap_error = "NULL"; return CLK_EINVAL; }
    PcapFile pf; int rc = pf.open(path, *cursor);
    if (rc != CLK_OK) return rc;
    uint64_t bytes = 0; bool eof = false;
    rc = pack(pf, blob, cap_bytes, offs16, cap_frames, 0, n_frames, &bytes, &eof);
    *cursor = eof ? 0 : (uint64_t)ftello(pf.f);                         // 0 = nothing left to read
    return rc;
} CLK_GUARD_INT(nullptr)

int clk_classify_pcap(clk_engine *e, const char *path, uint32_t flags, uint32_t *verdicts, uint64_t cap, uint64_t *n)
try {
    if (!e || !path || !n || (cap && !verdicts)) { tl_pcap_error = "NULL"; return CLK_EINVAL; }
    PcapFile pf; int rc = pf.open(path, 0);
    if (rc != CLK_OK) return rc;
    const uint64_t kBytes = 256ull << 20, kFrames = 2ull << 20;          // one chunk: what the host pipeline splits further
    struct Buf { uint8_t *blob = nullptr; uint32_t *offs = nullptr; uint64_t frames = 0, bytes = 0; bool eof = false; int rc = CLK_OK; } buf[2];
    auto release = [&]() { for (auto &b : buf) { if (b.blob) clk_host_free(b.blob); if (b.offs) clk_host_free(b.offs); } };
    for (auto &b : buf)
        if (clk_host_alloc(kBytes, (void **)&b.blob) != CLK_OK || clk_host_alloc((kFrames + 1) * 4, (void **)&b.offs) != CLK_OK) { release(); tl_pcap_error = "pinned host allocation failed"; return CLK_ENOMEM; }
    auto fill = [&](int i) { Buf &b = buf[i]; b.rc = pack(pf, b.blob, kBytes, b.offs, kFrames, 0, &b.frames, &b.bytes, &b.eof); };
    uint64_t total = 0; int cur = 0;
    fill(cur);
    for (;;) {
        Buf &b = buf[cur];
        if (b.rc != CLK_OK) { rc = b.rc; break; }
        if (b.frames == 0) {
            if (!b.eof) { tl_pcap_error = "a frame larger than the 256 MiB chunk"; rc = CLK_E2BIG; }
            break;
        }
        std::future<void> next;
        const bool more = !b.eof;
        if (more) next = std::async(std::launch::async, fill, cur ^ 1);     // read + pack chunk k+1 while chunk k is on the GPU
        const uint64_t take = total + b.frames <= cap ? b.frames : (cap > total ? cap - total : 0);
        if (take) rc = clk_classify_varlen(e, b.blob, b.offs, take, flags, verdicts + total);
        if (more) next.wait();
        if (rc != CLK_OK) { tl_pcap_error = clk_last_error(e); break; }
        total += b.frames;
        if (!more) break;
        cur ^= 1;
    }
    release();
    *n = total;                                                         // frames in the file (may exceed cap: verdicts holds the first cap)
    return rc;
} CLK_GUARD_INT(e)

}  // extern "C"
