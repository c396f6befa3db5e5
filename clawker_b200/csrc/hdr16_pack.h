// hdr16_pack.h — one frame -> one clk_hdr16 (include/clawker_gpu.h "compact ingest"); shared by clk_pack_hdr16 (engine.cpp, slot
// batches) and clk_afpacket_read_hdr16 (ingest.cpp, straight from the capture ring).
#pragma once
#include <cstdint>
#include <cstring>
#include "../../include/clawker_gpu.h"

// p: at least 64 readable bytes of the frame (zero padded if the capture was shorter).  Returns true when the record needs an aux slot
// (IPv6, IPv4 options, later fragments): the caller stores the 64 bytes and sets r.kind = CLK_HDR_AUX, r.src = slot index.
static inline bool clk_hdr16_from_frame(const uint8_t *p, clk_hdr16 &r)
{
    memset(&r, 0, sizeof r);
    const bool ipv4 = p[12] == 0x08 && p[13] == 0x00, ipv6 = p[12] == 0x86 && p[13] == 0xDD;
    if (ipv4 && p[14] == 0x45 && ((p[20] & 0x1f) | p[21]) == 0) {          // IHL 5, fragment offset 0 (same test as fast_classify)
        r.kind = CLK_HDR_IPV4; r.proto = p[23];
        memcpy(&r.src, p + 26, 4); memcpy(&r.dst, p + 30, 4);
        r.sport = (uint16_t)((p[34] << 8) | p[35]); r.dport = (uint16_t)((p[36] << 8) | p[37]);
        return false;
    }
    if (!ipv4 && !ipv6) { r.kind = CLK_HDR_NONE; return false; }
    return true;
}
