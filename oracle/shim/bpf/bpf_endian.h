/* TEST INFRASTRUCTURE — userspace stand-in for libbpf's <bpf/bpf_endian.h> (little-endian host). */
#ifndef CLK_SHIM_BPF_ENDIAN_H
#define CLK_SHIM_BPF_ENDIAN_H
#if __BYTE_ORDER__ != __ORDER_LITTLE_ENDIAN__
#error "oracle shim assumes a little-endian host"
#endif
#define bpf_htons(x) ((__u16)__builtin_bswap16((__u16)(x)))
#define bpf_ntohs(x) ((__u16)__builtin_bswap16((__u16)(x)))
#define bpf_htonl(x) ((__u32)__builtin_bswap32((__u32)(x)))
#define bpf_ntohl(x) ((__u32)__builtin_bswap32((__u32)(x)))
#endif
