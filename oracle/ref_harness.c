/*
 * ref_harness.c — runs the REFERENCE's own eBPF programs on the host.
 *
 * TEST INFRASTRUCTURE (parity oracle / CPU baseline of kind "reference").  Linked with
 * oracle/_ref/clawker_ref.o, which oracle/Makefile compiles straight from
 * /root/reference/internal/ebpf/bpf/clawker.c (+ common.h) against oracle/shim — the
 * reference sources are never copied into this repository.
 *
 * This file supplies what the kernel would: the five BPF helpers the programs call
 * (declared in oracle/shim/bpf/bpf_helpers.h) backed by the oracle's hash maps, and
 * ref_run_event(), which fills a struct bpf_sock_addr / bpf_sock exactly as the kernel's
 * cgroup hooks would and invokes the matching SEC("cgroup/...") program
 * (clawker.c:69,90,111,134,167,198,224).
 */
#include <stdint.h>
#include <string.h>
#include <linux/types.h>
#include <linux/bpf.h>
#include "../include/clawker_gpu.h"

typedef struct orc_tables orc_tables;
typedef struct orc_metrics orc_metrics;
void *orc__container_get(const orc_tables *t, const void *k);
void *orc__bypass_get(const orc_tables *t, const void *k);
void *orc__dns_get(const orc_tables *t, const void *k);
void *orc__route_get(const orc_tables *t, const void *k);
void *orc__metric_get(orc_metrics *m, const void *k);
void  orc__metric_put(orc_metrics *m, const void *k, const void *v);

/* the reference's pinned maps (common.h:121-177) — only their addresses are used */
extern char container_map[], bypass_map[], dns_cache[], route_map[], metrics_map[];
/* the reference's programs */
int clawker_connect4(struct bpf_sock_addr *ctx);
int clawker_sendmsg4(struct bpf_sock_addr *ctx);
int clawker_recvmsg4(struct bpf_sock_addr *ctx);
int clawker_connect6(struct bpf_sock_addr *ctx);
int clawker_sendmsg6(struct bpf_sock_addr *ctx);
int clawker_recvmsg6(struct bpf_sock_addr *ctx);
int clawker_sock_create(struct bpf_sock *ctx);

struct ref_out { int ret; uint32_t ip; uint16_t port; int metric_action; };

static __thread struct {
    const orc_tables *t; orc_metrics *mx; const clk_event *ev; int metric_action; uint64_t sink;
} cur;

void *bpf_map_lookup_elem(void *map, const void *key)
{
    if (map == (void *)container_map) return orc__container_get(cur.t, key);
    if (map == (void *)bypass_map)    return orc__bypass_get(cur.t, key);
    if (map == (void *)dns_cache)     return orc__dns_get(cur.t, key);
    if (map == (void *)route_map)     return orc__route_get(cur.t, key);
    if (map == (void *)metrics_map) {
        cur.metric_action = ((const clk_metric_key *)key)->action;
        if (!cur.mx) { cur.sink = 0; return &cur.sink; }
        return orc__metric_get(cur.mx, key);
    }
    return NULL;
}
long bpf_map_update_elem(void *map, const void *key, const void *value, __u64 flags)
{
    (void)flags;
    if (map == (void *)metrics_map) { orc__metric_put(cur.mx, key, value); return 0; }
    return -1;
}
__u64 bpf_get_current_uid_gid(void)   { return (__u64)cur.ev->uid; }
__u64 bpf_get_current_cgroup_id(void) { return cur.ev->cgroup_id; }
long bpf_getsockopt(void *ctx, int level, int optname, void *optval, int optlen)
{
    (void)ctx;
    if (level == 1 && optname == 36 && optlen == 4) { memcpy(optval, &cur.ev->so_mark, 4); return 0; }
    return -1;
}

void ref_run_event(const orc_tables *t, orc_metrics *mx, const clk_event *ev, struct ref_out *out)
{
    cur.t = t; cur.mx = mx; cur.ev = ev; cur.metric_action = -1;
    struct bpf_sock_addr ctx; memset(&ctx, 0, sizeof ctx);
    int v6 = ev->hook >= CLK_HOOK_CONNECT6 && ev->hook <= CLK_HOOK_RECVMSG6;
    ctx.user_family = ctx.family = v6 ? 10 : 2;
    ctx.type = ev->sock_type;
    ctx.user_ip4 = ev->user_ip4;
    memcpy(ctx.user_ip6, ev->user_ip6, 16);
    ctx.user_port = (uint16_t)__builtin_bswap16(ev->user_port);
    int ret = 1;
    switch (ev->hook) {
    case CLK_HOOK_CONNECT4: ret = clawker_connect4(&ctx); break;
    case CLK_HOOK_SENDMSG4: ret = clawker_sendmsg4(&ctx); break;
    case CLK_HOOK_RECVMSG4: ret = clawker_recvmsg4(&ctx); break;
    case CLK_HOOK_CONNECT6: ret = clawker_connect6(&ctx); break;
    case CLK_HOOK_SENDMSG6: ret = clawker_sendmsg6(&ctx); break;
    case CLK_HOOK_RECVMSG6: ret = clawker_recvmsg6(&ctx); break;
    case CLK_HOOK_SOCK_CREATE: {
        struct bpf_sock sk; memset(&sk, 0, sizeof sk); sk.type = ev->sock_type; sk.family = 2;
        ret = clawker_sock_create(&sk); break;
    }
    default: break;
    }
    out->ret = ret;
    out->ip = v6 ? ctx.user_ip6[3] : ctx.user_ip4;
    out->port = (uint16_t)__builtin_bswap16((uint16_t)(ctx.user_port & 0xffff));
    out->metric_action = cur.metric_action;
}
