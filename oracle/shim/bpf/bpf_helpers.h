/*
 * TEST INFRASTRUCTURE — not part of the product (see oracle/README.md).
 *
 * Userspace stand-in for libbpf's <bpf/bpf_helpers.h>, just enough to compile the
 * reference's internal/ebpf/bpf/clawker.c + common.h for the host with gcc so the
 * reference's own decision code can be executed as the parity oracle.
 * The helpers the reference calls (common.h:199,208,259,263,266,273,333,387,394) are
 * declared here as ordinary functions; oracle/ref_harness.c implements them.
 */
#ifndef CLK_SHIM_BPF_HELPERS_H
#define CLK_SHIM_BPF_HELPERS_H

#include <linux/types.h>

#undef __always_inline
#define __always_inline inline __attribute__((always_inline))

#define SEC(name) __attribute__((used))
#define __uint(name, val) int (*name)[val]
#define __type(name, val) __typeof__(val) *name
#define LIBBPF_PIN_BY_NAME 1

void *bpf_map_lookup_elem(void *map, const void *key);
long  bpf_map_update_elem(void *map, const void *key, const void *value, __u64 flags);
__u64 bpf_get_current_uid_gid(void);
__u64 bpf_get_current_cgroup_id(void);
long  bpf_getsockopt(void *ctx, int level, int optname, void *optval, int optlen);

#endif
