/*
 * clk_oracle.c — CPU oracle for the clawker firewall verdict path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load this; the product (clawker_b200/, libclawker_gpu.so) never
 * does.  It exists to define, on the CPU, the exact verdict word the CUDA path must
 * produce for every packet.
 *
 * Two builds of this one file (oracle/Makefile):
 *   liboracle_port.so        decision stage = the C restatement below ("port")
 *   _ref/liboracle_ref.so    decision stage = the reference's own eBPF C
 *                            (/root/reference/internal/ebpf/bpf/clawker.c + common.h,
 *                            compiled for the host against oracle/shim, never copied)
 * Everything around the decision stage (packet -> socket-event mapping, name stage,
 * verdict word encoding) is shared, because the reference has no counterpart for it:
 * it is the mapping defined in SURVEY.md §8(d) / DESIGN.md §3.
 *
 * Parity status of each stage:
 *   decide_*  (A1-A5,A9)   pinned: port == _ref on exhaustive/random events (tests/test_oracle_ref.py)
 *   DomainHash (A7)        pinned by the FNV-1a-32 definition + published FNV vectors
 *   zone select (A13)      parity unpinned: restates coredns v1.14.2 core/dnsserver/server.go ServeDNS
 *   SNI match  (A14)       parity unpinned: restates Envoy v1.37 filter-chain server_names matching
 *   DNS answers (A10)      restates internal/dnsbpf/dnsbpf.go:75-97 over an RFC 1035 walk
 *                          (miekg/dns v1.1.72 not in tree: wire parsing parity unpinned)
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <stdio.h>
#include <pthread.h>
#include <sched.h>

#include "../include/clawker_gpu.h"
#include "hmap.h"

/* ------------------------------------------------------------------------------------------------
 * tables
 * ---------------------------------------------------------------------------------------------- */

typedef struct strtab {           /* string -> u32 value, exact compare */
    char   **s;
    uint32_t *len, *val, *next;
    uint32_t n, cap;
    hmap     idx;                 /* u64 hash -> u32 head index + 1 */
} strtab;

typedef struct maskset {          /* one distinct CIDR mask and its network set */
    uint32_t mask;
    hmap     nets;                /* u32 addr -> u8 */
} maskset;

typedef struct http_vhost {       /* one Envoy virtual host of the raw_buffer chain (envoy.go:403-470) */
    char **domains; uint32_t n_domains;                       /* httpDomains(), envoy.go:932-944 */
    char **prefixes; uint8_t *allow; uint32_t n_routes;       /* buildHTTPRoutes(), envoy.go:566-607: first prefix match wins */
} http_vhost;

typedef struct orc_tables {
    hmap container, bypass, dns, route, endpoint, endpoint6;
    clk_cidr *cidrs; uint32_t n_cidrs, cap_cidrs;
    maskset  *masks; uint32_t n_masks;
    strtab zones, snis;
    struct http_vhost *vhosts; uint32_t n_vhosts;   /* http_egress_routes virtual hosts, in config order (deny_all is implicit) */
    uint32_t coredns_ips[8]; uint32_t n_coredns;
} orc_tables;

typedef struct orc_metrics { hmap m; } orc_metrics;   /* clk_metric_key -> u64 (one per thread: PERCPU) */

static uint64_t fnv64(const uint8_t *p, uint32_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint32_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
static void strtab_init(strtab *t) { memset(t, 0, sizeof *t); hmap_init(&t->idx, 8, 4, 64); }
static void strtab_free(strtab *t)
{
    for (uint32_t i = 0; i < t->n; i++) free(t->s[i]);
    free(t->s); free(t->len); free(t->val); free(t->next); hmap_free(&t->idx);
}
static int strtab_find(const strtab *t, const uint8_t *p, uint32_t n, uint32_t *val)
{
    uint64_t h = fnv64(p, n);
    uint32_t *head = (uint32_t *)hmap_get(&t->idx, &h);
    if (!head) return 0;
    for (uint32_t i = *head; i; i = t->next[i - 1])
        if (t->len[i - 1] == n && !memcmp(t->s[i - 1], p, n)) { if (val) *val = t->val[i - 1]; return 1; }
    return 0;
}
static void strtab_add(strtab *t, const char *s, uint32_t val)
{
    uint32_t n = (uint32_t)strlen(s);
    if (strtab_find(t, (const uint8_t *)s, n, NULL)) return;
    if (t->n == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 64;
        t->s = realloc(t->s, t->cap * sizeof *t->s); t->len = realloc(t->len, t->cap * 4);
        t->val = realloc(t->val, t->cap * 4); t->next = realloc(t->next, t->cap * 4);
    }
    uint64_t h = fnv64((const uint8_t *)s, n);
    uint32_t *head = (uint32_t *)hmap_get(&t->idx, &h);
    t->s[t->n] = strdup(s); t->len[t->n] = n; t->val[t->n] = val; t->next[t->n] = head ? *head : 0;
    uint32_t id = ++t->n;
    hmap_put(&t->idx, &h, &id);
}

/* ebpf.DomainHash, internal/ebpf/types.go:121-125: FNV-1a-32 over strings.ToLower(domain). */
uint32_t orc_domain_hash(const char *s, uint32_t n)
{
    uint32_t h = 0x811c9dc5u;
    for (uint32_t i = 0; i < n; i++) {
        uint8_t c = (uint8_t)s[i];
        if (c >= 'A' && c <= 'Z') c += 32;
        h ^= c; h *= 0x01000193u;
    }
    return h;
}

orc_tables *orc_tables_new(void)
{
    orc_tables *t = calloc(1, sizeof *t);
    hmap_init(&t->container, 8, sizeof(clk_container_config), 256);
    hmap_init(&t->bypass, 8, 1, 256);
    hmap_init(&t->dns, 4, sizeof(clk_dns_entry), 16384);
    hmap_init(&t->route, sizeof(clk_route_key), sizeof(clk_route_val), 8192);
    hmap_init(&t->endpoint, 4, sizeof(clk_endpoint), 256);
    hmap_init(&t->endpoint6, 16, sizeof(clk_endpoint), 256);
    strtab_init(&t->zones); strtab_init(&t->snis);
    return t;
}
void orc_tables_free(orc_tables *t)
{
    hmap_free(&t->container); hmap_free(&t->bypass); hmap_free(&t->dns); hmap_free(&t->route);
    hmap_free(&t->endpoint); hmap_free(&t->endpoint6);
    for (uint32_t i = 0; i < t->n_masks; i++) hmap_free(&t->masks[i].nets);
    for (uint32_t i = 0; i < t->n_vhosts; i++) {
        http_vhost *v = &t->vhosts[i];
        for (uint32_t k = 0; k < v->n_domains; k++) free(v->domains[k]);
        for (uint32_t k = 0; k < v->n_routes; k++) free(v->prefixes[k]);
        free(v->domains); free(v->prefixes); free(v->allow);
    }
    free(t->vhosts);
    free(t->masks); free(t->cidrs); strtab_free(&t->zones); strtab_free(&t->snis); free(t);
}
void orc_set_container(orc_tables *t, uint64_t cg, const clk_container_config *c)
{
    hmap_put(&t->container, &cg, c);
    for (uint32_t i = 0; i < t->n_coredns; i++) if (t->coredns_ips[i] == c->coredns_ip) return;
    if (t->n_coredns < 8) t->coredns_ips[t->n_coredns++] = c->coredns_ip;
}
void orc_set_bypass(orc_tables *t, uint64_t cg, uint8_t v) { hmap_put(&t->bypass, &cg, &v); }
void orc_set_dns(orc_tables *t, uint32_t ip, uint32_t hash, uint32_t expire)
{
    clk_dns_entry e = { hash, expire }; hmap_put(&t->dns, &ip, &e);
}
void orc_set_route(orc_tables *t, uint32_t hash, uint16_t port, uint16_t envoy_port)
{
    clk_route_key k = { hash, port, 0 }; clk_route_val v = { envoy_port, 0 }; hmap_put(&t->route, &k, &v);
}
void orc_set_endpoint(orc_tables *t, uint32_t ip, const clk_endpoint *e) { hmap_put(&t->endpoint, &ip, e); }
void orc_set_endpoint6(orc_tables *t, const uint8_t ip6[16], const clk_endpoint *e) { hmap_put(&t->endpoint6, ip6, e); }
void orc_add_cidr(orc_tables *t, uint32_t addr, uint32_t mask)
{
    if (t->n_cidrs == t->cap_cidrs) { t->cap_cidrs = t->cap_cidrs ? t->cap_cidrs * 2 : 64; t->cidrs = realloc(t->cidrs, t->cap_cidrs * sizeof *t->cidrs); }
    t->cidrs[t->n_cidrs].addr = addr; t->cidrs[t->n_cidrs].mask = mask; t->n_cidrs++;
    uint32_t i; uint8_t one = 1; uint32_t net = addr & mask;
    for (i = 0; i < t->n_masks; i++) if (t->masks[i].mask == mask) break;
    if (i == t->n_masks) {
        t->masks = realloc(t->masks, (t->n_masks + 1) * sizeof *t->masks);
        t->masks[i].mask = mask; hmap_init(&t->masks[i].nets, 4, 1, 64); t->n_masks++;
    }
    /* is_in_subnet (common.h:220-223) is (ip & mask) == addr: a rule whose addr has host bits
     * set can never match; keep that exactly by only indexing canonical networks. */
    if (net == addr) hmap_put(&t->masks[i].nets, &net, &one);
}
/* zone: normalizeDomain()'d Corefile zone (coredns.go:58); value = DomainHash(zone). */
void orc_add_zone(orc_tables *t, const char *zone) { strtab_add(&t->zones, zone, orc_domain_hash(zone, (uint32_t)strlen(zone))); }
/* server_names entry as emitted by serverNames() envoy.go:905-915 ("d" or ".d"). */
void orc_add_sni(orc_tables *t, const char *name) { strtab_add(&t->snis, name, t->snis.n); }
/* One virtual host: domains and routes as '\n'-joined lists; a route is 'A' or 'D' followed by its path prefix. */
void orc_add_http_vhost(orc_tables *t, const char *domains, const char *routes)
{
    t->vhosts = realloc(t->vhosts, (t->n_vhosts + 1) * sizeof *t->vhosts);
    http_vhost *v = &t->vhosts[t->n_vhosts++]; memset(v, 0, sizeof *v);
    for (const char *p = domains; *p;) {
        const char *e = strchr(p, '\n'); size_t n = e ? (size_t)(e - p) : strlen(p);
        v->domains = realloc(v->domains, (v->n_domains + 1) * sizeof(char *));
        v->domains[v->n_domains] = calloc(1, n + 1); memcpy(v->domains[v->n_domains++], p, n);
        p += n + (e ? 1 : 0);
    }
    for (const char *p = routes; *p;) {
        const char *e = strchr(p, '\n'); size_t n = e ? (size_t)(e - p) : strlen(p);
        v->prefixes = realloc(v->prefixes, (v->n_routes + 1) * sizeof(char *));
        v->allow = realloc(v->allow, v->n_routes + 1);
        v->allow[v->n_routes] = p[0] == 'A';
        v->prefixes[v->n_routes] = calloc(1, n); memcpy(v->prefixes[v->n_routes++], p + 1, n - 1);
        p += n + (e ? 1 : 0);
    }
}

orc_metrics *orc_metrics_new(void) { orc_metrics *m = calloc(1, sizeof *m); hmap_init(&m->m, sizeof(clk_metric_key), 8, 1024); return m; }
void orc_metrics_free(orc_metrics *m) { hmap_free(&m->m); free(m); }
uint32_t orc_metrics_count(const orc_metrics *m) { return m->m.used; }
uint32_t orc_metrics_dump(const orc_metrics *m, clk_metric *out, uint32_t cap)
{
    uint32_t n = 0;
    for (uint32_t i = 0; i < m->m.cap && n < cap; i++)
        if (m->m.occ[i]) { memcpy(&out[n].key, m->m.keys + (size_t)i * 16, 16); memcpy(&out[n].count, m->m.vals + (size_t)i * 8, 8); n++; }
    return n;
}
static void metrics_add(orc_metrics *m, const clk_metric_key *k, uint64_t by)
{
    uint64_t *v = (uint64_t *)hmap_get(&m->m, k);
    if (v) *v += by; else hmap_put(&m->m, k, &by);
}
void orc_metrics_merge(orc_metrics *dst, const orc_metrics *src)
{
    for (uint32_t i = 0; i < src->m.cap; i++)
        if (src->m.occ[i]) { uint64_t c; memcpy(&c, src->m.vals + (size_t)i * 8, 8); metrics_add(dst, (const clk_metric_key *)(src->m.keys + (size_t)i * 16), c); }
}

/* ------------------------------------------------------------------------------------------------
 * decision stage — result in observable form
 * ---------------------------------------------------------------------------------------------- */

typedef struct decision {
    uint8_t  verdict;      /* CLK_V_*  */
    uint8_t  metric;       /* CLK_M_*  */
    uint32_t new_ip;       /* valid when verdict == REWRITE (NBO-in-native)  */
    uint16_t new_port;     /* host order                                      */
} decision;

#define SOCK_STREAM_ 1
#define SOCK_DGRAM_  2
#define SOCK_RAW_    3
#define CLAWKER_MARK_ 0xC1A4u
#define DOCKER_DNS_NBO 0x0b00007fu      /* bpf_htonl(0x7f00000b) on little-endian, common.h:62 */

#ifndef CLK_ORACLE_USE_REF
/* ---- the restatement ("port") -------------------------------------------------------------- */

static void emit_metric(orc_metrics *mx, uint64_t cg, uint32_t hash, uint16_t port, uint8_t action, decision *d)
{
    clk_metric_key k; memset(&k, 0, sizeof k);
    k.cgroup_id = cg; k.domain_hash = hash; k.dst_port = port; k.action = action;
    if (mx) metrics_add(mx, &k, 1);
    d->metric = (uint8_t)(action + 1);
}


static int is_loopback4(uint32_t ip) { return (ip & 0x000000ffu) == 0x0000007fu; }   /* common.h:213-216 on LE */

/* A12 (extension): dst inside any allowed CIDR, evaluated as is_in_subnet over all rules. */
static int cidr_hit(const orc_tables *t, uint32_t ip, int linear)
{
    if (linear) {
        for (uint32_t i = 0; i < t->n_cidrs; i++) if ((ip & t->cidrs[i].mask) == t->cidrs[i].addr) return 1;
        return 0;
    }
    for (uint32_t i = 0; i < t->n_masks; i++) { uint32_t net = ip & t->masks[i].mask; if (hmap_get(&t->masks[i].nets, &net)) return 1; }
    return 0;
}

/* decide_connect, common.h:323-410 */
static void decide_connect(const orc_tables *t, orc_metrics *mx, const clk_container_config *cfg, uint64_t cg,
                           uint32_t mark, uint8_t type, uint32_t dst_ip, uint16_t dst_port, int linear, decision *d)
{
    if (mark == CLAWKER_MARK_) return;                                               /* :332-335 */
    if (dst_port == 53) {                                                            /* :339-345 */
        d->verdict = CLK_V_REWRITE; d->new_ip = cfg->coredns_ip; d->new_port = 53;
        emit_metric(mx, cg, 0, 53, CLK_ACTION_ALLOW, d); return;
    }
    if (is_loopback4(dst_ip)) return;                                                /* :347 */
    if ((dst_ip & cfg->net_mask) == cfg->net_addr) return;                           /* :350 */
    if (cidr_hit(t, dst_ip, linear)) return;                                         /* A12, at :350 */
    if (cfg->host_proxy_ip != 0 && dst_ip == cfg->host_proxy_ip && dst_port == cfg->host_proxy_port) return; /* :353 */
    if (dst_ip == cfg->gateway_ip) {                                                 /* :362-370 */
        if (cfg->host_proxy_port != 0 && dst_port == cfg->host_proxy_port) return;
        d->verdict = CLK_V_REWRITE; d->new_ip = cfg->envoy_ip; d->new_port = cfg->egress_port;
        emit_metric(mx, cg, 0, dst_port, CLK_ACTION_DENY, d); return;
    }
    if (type == SOCK_DGRAM_) {                                                       /* :374-378 */
        emit_metric(mx, cg, 0, dst_port, CLK_ACTION_DENY, d); d->verdict = CLK_V_DENY; return;
    }
    uint32_t domain_hash = 0;
    const clk_dns_entry *dns = (const clk_dns_entry *)hmap_get(&t->dns, &dst_ip);    /* :387 */
    if (dns) {
        domain_hash = dns->domain_hash;
        clk_route_key rk = { dns->domain_hash, dst_port, 0 };
        const clk_route_val *rv = (const clk_route_val *)hmap_get(&t->route, &rk);   /* :394 */
        if (rv) {
            d->verdict = CLK_V_REWRITE; d->new_ip = cfg->envoy_ip; d->new_port = rv->envoy_port;
            emit_metric(mx, cg, dns->domain_hash, dst_port, CLK_ACTION_ALLOW, d); return;
        }
    }
    d->verdict = CLK_V_REWRITE; d->new_ip = cfg->envoy_ip; d->new_port = cfg->egress_port;  /* :405-409 */
    emit_metric(mx, cg, domain_hash, dst_port, CLK_ACTION_ALLOW, d);
}

/* decide_sendmsg, common.h:416-439 */
static void decide_sendmsg(const orc_tables *t, orc_metrics *mx, const clk_container_config *cfg, uint64_t cg,
                           uint32_t dst_ip, uint16_t dst_port, int linear, decision *d)
{
    if (dst_port == 53) { d->verdict = CLK_V_REWRITE; d->new_ip = cfg->coredns_ip; d->new_port = 53; return; }
    if (is_loopback4(dst_ip)) return;
    if ((dst_ip & cfg->net_mask) == cfg->net_addr) return;
    if (cidr_hit(t, dst_ip, linear)) return;                                         /* A12 */
    emit_metric(mx, cg, 0, dst_port, CLK_ACTION_DENY, d); d->verdict = CLK_V_DENY;
}

/* enter_enforced, common.h:256-280.  Returns cfg or NULL for fast pass. */
static const clk_container_config *enter_enforced(const orc_tables *t, orc_metrics *mx, const clk_event *ev,
                                                  int check_bypass, decision *d)
{
    if (ev->uid == 0) return NULL;
    if (check_bypass) {
        const uint8_t *b = (const uint8_t *)hmap_get(&t->bypass, &ev->cgroup_id);
        if (b && *b == 1) { emit_metric(mx, ev->cgroup_id, 0, 0, CLK_ACTION_BYPASS, d); return NULL; }
    }
    return (const clk_container_config *)hmap_get(&t->container, &ev->cgroup_id);
}

/* the seven programs, clawker.c:69-240.  ip_io/port_io carry ctx->user_ip4|user_ip6[3] / user_port. */
static void run_hooks(const orc_tables *t, orc_metrics *mx, const clk_event *ev, int linear, decision *d)
{
    const clk_container_config *cfg;
    switch (ev->hook) {
    case CLK_HOOK_CONNECT4:
        if (ev->sock_type != SOCK_STREAM_ && ev->sock_type != SOCK_DGRAM_) return;   /* clawker.c:72 */
        if (!(cfg = enter_enforced(t, mx, ev, 1, d))) return;
        decide_connect(t, mx, cfg, ev->cgroup_id, ev->so_mark, ev->sock_type, ev->user_ip4, ev->user_port, linear, d);
        return;
    case CLK_HOOK_SENDMSG4:
        if (!(cfg = enter_enforced(t, mx, ev, 1, d))) return;
        decide_sendmsg(t, mx, cfg, ev->cgroup_id, ev->user_ip4, ev->user_port, linear, d);
        return;
    case CLK_HOOK_RECVMSG4:
        if (!(cfg = enter_enforced(t, mx, ev, 0, d))) return;
        if (ev->user_ip4 == cfg->coredns_ip && ev->user_port == 53) {                /* common.h:447-451 */
            d->verdict = CLK_V_REWRITE; d->new_ip = DOCKER_DNS_NBO; d->new_port = 53;
        }
        return;
    case CLK_HOOK_CONNECT6:
    case CLK_HOOK_SENDMSG6: {
        if (!(cfg = enter_enforced(t, mx, ev, 1, d))) return;
        const uint32_t *a = ev->user_ip6;
        if (a[0] == 0 && a[1] == 0 && a[2] == 0 && a[3] == 0x01000000u) return;      /* ::1, common.h:226-230 */
        if (a[0] == 0 && a[1] == 0 && a[2] == 0xffff0000u) {                         /* v4-mapped, :233-237 */
            if (ev->hook == CLK_HOOK_CONNECT6) {
                if (ev->sock_type != SOCK_STREAM_ && ev->sock_type != SOCK_DGRAM_) return;   /* clawker.c:146 */
                decide_connect(t, mx, cfg, ev->cgroup_id, ev->so_mark, ev->sock_type, a[3], ev->user_port, linear, d);
            } else
                decide_sendmsg(t, mx, cfg, ev->cgroup_id, a[3], ev->user_port, linear, d);
            return;
        }
        emit_metric(mx, ev->cgroup_id, 0, ev->user_port, CLK_ACTION_DENY, d);       /* clawker.c:155,186 */
        d->verdict = CLK_V_DENY;
        return;
    }
    case CLK_HOOK_RECVMSG6: {
        if (!(cfg = enter_enforced(t, mx, ev, 0, d))) return;
        const uint32_t *a = ev->user_ip6;
        if (!(a[0] == 0 && a[1] == 0 && a[2] == 0xffff0000u)) return;
        if (a[3] == cfg->coredns_ip && ev->user_port == 53) { d->verdict = CLK_V_REWRITE; d->new_ip = DOCKER_DNS_NBO; d->new_port = 53; }
        return;
    }
    case CLK_HOOK_SOCK_CREATE:
        if (!(cfg = enter_enforced(t, mx, ev, 1, d))) return;
        if (ev->sock_type == SOCK_RAW_) { emit_metric(mx, ev->cgroup_id, 0, 0, CLK_ACTION_DENY, d); d->verdict = CLK_V_DENY; }
        return;
    default: return;
    }
}

const char *orc_kind(void) { return "port"; }

#else  /* CLK_ORACLE_USE_REF ------------------------------------------------------------------- */

/* Implemented in oracle/ref_harness.c, which links the reference's clawker.c object. */
struct ref_out { int ret; uint32_t ip; uint16_t port; int metric_action; };
void ref_run_event(const orc_tables *t, orc_metrics *mx, const clk_event *ev, struct ref_out *out);

static void run_hooks(const orc_tables *t, orc_metrics *mx, const clk_event *ev, int linear, decision *d)
{
    (void)linear;
    struct ref_out o;
    uint32_t ip_in = (ev->hook >= CLK_HOOK_CONNECT6 && ev->hook <= CLK_HOOK_RECVMSG6) ? ev->user_ip6[3] : ev->user_ip4;
    ref_run_event(t, mx, ev, &o);
    if (o.metric_action >= 0) d->metric = (uint8_t)(o.metric_action + 1);
    if (o.ret == 0) { d->verdict = CLK_V_DENY; return; }
    if (ev->hook == CLK_HOOK_SOCK_CREATE || ev->hook == CLK_HOOK_NONE) return;
    if (o.ip != ip_in || o.port != ev->user_port) { d->verdict = CLK_V_REWRITE; d->new_ip = o.ip; d->new_port = o.port; }
}
const char *orc_kind(void) { return "reference"; }
/* accessors for the harness */
void *orc__container_get(const orc_tables *t, const void *k) { return hmap_get(&t->container, k); }
void *orc__bypass_get(const orc_tables *t, const void *k) { return hmap_get(&t->bypass, k); }
void *orc__dns_get(const orc_tables *t, const void *k) { return hmap_get(&t->dns, k); }
void *orc__route_get(const orc_tables *t, const void *k) { return hmap_get(&t->route, k); }
void *orc__metric_get(orc_metrics *m, const void *k) { return m ? hmap_get(&m->m, k) : NULL; }
void  orc__metric_put(orc_metrics *m, const void *k, const void *v) { if (m && !hmap_get(&m->m, k)) hmap_put(&m->m, k, v); }
#endif

/* ------------------------------------------------------------------------------------------------
 * name stage (shared): DNS question -> zone, TLS ClientHello -> SNI match, DNS answers -> updates
 * ---------------------------------------------------------------------------------------------- */

static inline uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }
static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* Unpack a (possibly compressed) domain name from DNS message m[0..mlen) at *off.
 * Emits the lower-cased text form into buf (labels joined with '.', no trailing dot),
 * label start offsets into starts[], and the index of the last label that contains a
 * byte outside [a-z0-9_-] (such a label can never be part of a zone)
 * into *last_bad (-1 none).  Advances *off past the name as stored.  Returns the number
 * of labels, or -1 on malformed input (RFC 1035 §4.1.4; miekg/dns UnpackDomainName limits:
 * 255 wire octets, 126 pointers). */
static int dns_unpack_name(const uint8_t *m, uint32_t mlen, uint32_t *off, uint8_t *buf, uint32_t *blen,
                           uint16_t *starts, int *last_bad)
{
    uint32_t p = *off, wire = 0, n = 0; int labels = 0, ptrs = 0, jumped = 0;
    *last_bad = -1;
    for (;;) {
        if (p >= mlen) return -1;
        uint32_t c = m[p];
        if ((c & 0xC0) == 0xC0) {
            if (p + 1 >= mlen) return -1;
            if (!jumped) { *off = p + 2; jumped = 1; }
            if (++ptrs > 126) return -1;
            p = ((c & 0x3F) << 8) | m[p + 1];
            continue;
        }
        if (c & 0xC0) return -1;
        if (c == 0) { if (!jumped) *off = p + 1; break; }
        if (p + 1 + c > mlen) return -1;
        wire += c + 1; if (wire + 1 > 255) return -1;
        if (labels) buf[n++] = '.';
        if (labels < 128) starts[labels] = (uint16_t)n;
        for (uint32_t i = 0; i < c; i++) {
            uint8_t b = m[p + 1 + i];
            if (b >= 'A' && b <= 'Z') b += 32;
            if (!((b >= 'a' && b <= 'z') || (b >= '0' && b <= '9') || b == '_' || b == '-')) *last_bad = labels;
            buf[n++] = b;
        }
        labels++;
        p += 1 + c;
    }
    *blen = n;
    return labels;
}

/* CoreDNS zone selection (coredns v1.14.2 core/dnsserver/server.go ServeDNS): longest label-suffix
 * of the lower-cased qname that is a configured zone; "." otherwise.  Returns 1 + zone hash on a
 * non-root match, 0 for the root zone (template NXDOMAIN, coredns.go:86-93). */
static int zone_select(const orc_tables *t, const uint8_t *name, uint32_t nlen, const uint16_t *starts, int labels,
                       int last_bad, uint32_t *zone_hash)
{
    for (int k = last_bad + 1; k < labels; k++) {
        uint32_t s = starts[k];
        if (strtab_find(&t->zones, name + s, nlen - s, zone_hash)) return 1;
    }
    return 0;
}

/* DNS query sent to CoreDNS.  miekg/dns defaultMsgAcceptFunc + CoreDNS ServeDNS. */
static uint32_t dns_query_verdict(const orc_tables *t, const uint8_t *m, uint32_t mlen)
{
    if (mlen < 12) return CLK_N_DENY;
    uint32_t flags = be16(m + 2);
    if (flags & 0x8000) return CLK_N_DENY;                       /* responses are ignored          */
    uint32_t opcode = (flags >> 11) & 15;
    if (opcode != 0 && opcode != 4) return CLK_N_DENY;           /* NOTIMP                         */
    if (be16(m + 4) != 1) return CLK_N_DENY;                     /* qdcount != 1 -> FORMERR        */
    if (be16(m + 6) > 1 || be16(m + 8) > 1 || be16(m + 10) > 2) return CLK_N_DENY;
    uint8_t buf[260]; uint16_t starts[128]; uint32_t blen, off = 12; int last_bad;
    int labels = dns_unpack_name(m, mlen, &off, buf, &blen, starts, &last_bad);
    if (labels < 0 || off + 4 > mlen) return CLK_N_DENY;
    if (be16(m + off + 2) != 1) return CLK_N_DENY;               /* template IN ANY: non-IN class never answered */
    uint32_t zh;
    return zone_select(t, buf, blen, starts, labels, last_bad, &zh) ? CLK_N_ALLOW : CLK_N_DENY;
}

/* Skip one RR at *off: name, type(2) class(2) ttl(4) rdlen(2) rdata.  Returns 0 ok / -1 malformed. */
static int dns_skip_name(const uint8_t *m, uint32_t mlen, uint32_t *off)
{
    uint32_t p = *off, wire = 0;
    for (;;) {
        if (p >= mlen) return -1;
        uint32_t c = m[p];
        if ((c & 0xC0) == 0xC0) { if (p + 1 >= mlen) return -1; *off = p + 2; return 0; }
        if (c & 0xC0) return -1;
        if (c == 0) { *off = p + 1; return 0; }
        wire += c + 1; if (wire + 1 > 255) return -1;
        p += 1 + c;
    }
}

/* dnsbpf.Handler.writeARecords, internal/dnsbpf/dnsbpf.go:75-97, over the wire form.
 * Returns number of updates (<= CLK_MAX_A_PER_PACKET written to out when out != NULL);
 * *name_v = ALLOW when the question's zone is a dnsbpf zone, else DENY. */
static uint32_t dns_answer_updates(const orc_tables *t, const uint8_t *m, uint32_t mlen, uint32_t pkt_index,
                                   clk_dns_update_rec *out, uint32_t *name_v)
{
    *name_v = CLK_N_DENY;
    if (mlen < 12) return 0;
    uint32_t flags = be16(m + 2);
    if (!(flags & 0x8000)) return 0;                             /* not a response                 */
    if (be16(m + 4) != 1) return 0;
    uint32_t an = be16(m + 6), rest = be16(m + 8) + be16(m + 10);
    uint8_t buf[260]; uint16_t starts[128]; uint32_t blen, off = 12; int last_bad;
    int labels = dns_unpack_name(m, mlen, &off, buf, &blen, starts, &last_bad);
    if (labels < 0 || off + 4 > mlen) return 0;
    off += 4;
    uint32_t zone_hash;
    if (!zone_select(t, buf, blen, starts, labels, last_bad, &zone_hash)) return 0;   /* "." has no dnsbpf */
    *name_v = CLK_N_ALLOW;
    /* framing pass first: miekg fails the whole message on any malformed RR (no partial writes) */
    uint32_t p = off;
    for (uint32_t i = 0; i < an + rest; i++) {
        if (dns_skip_name(m, mlen, &p)) return 0;
        if (p + 10 > mlen) return 0;
        uint32_t type = be16(m + p), rdlen = be16(m + p + 8);
        if (p + 10 + rdlen > mlen) return 0;
        if (i < an && type == 1 && rdlen != 4) return 0;
        p += 10 + rdlen;
    }
    uint32_t n = 0; p = off;
    for (uint32_t i = 0; i < an; i++) {
        dns_skip_name(m, mlen, &p);
        uint32_t type = be16(m + p), ttl = ((uint32_t)be16(m + p + 4) << 16) | be16(m + p + 6), rdlen = be16(m + p + 8);
        if (type == 1) {                                         /* *dns.A, dnsbpf.go:83-86        */
            uint32_t ip = ld32(m + p + 10);                      /* IPToUint32, types.go:86-92     */
            if (ip != 0 && n < CLK_MAX_A_PER_PACKET) {           /* dnsbpf.go:88-90                */
                if (out) { out[n].ip = ip; out[n].domain_hash = zone_hash; out[n].ttl = ttl ? ttl : 60; out[n].packet = pkt_index; }
                n++;
            }
        }
        p += 10 + rdlen;
    }
    return n;
}

/* Envoy filter-chain server_names lookup (FilterChainManagerImpl::findFilterChainForServerName):
 * exact name, then ".suffix" for each '.' at index >= 1 (longest first) while pos < len - 1. */
static int sni_match(const orc_tables *t, const uint8_t *name, uint32_t n)
{
    if (strtab_find(&t->snis, name, n, NULL)) return 1;
    for (uint32_t pos = 1; pos + 1 < n; pos++)
        if (name[pos] == '.' && strtab_find(&t->snis, name + pos, n - pos, NULL)) return 1;
    return 0;
}

/* TLS ClientHello in one TCP segment -> SNI -> allow/deny.  A complete ClientHello that is malformed or does not
 * carry exactly one listed host_name falls to the deny chain (envoy.go:648-664).  A ClientHello that does NOT fit
 * the segment (record or handshake message longer than the captured payload: large / post-quantum key shares)
 * cannot be judged from this packet — Envoy's tls_inspector keeps buffering — so it is N/A ("needs reassembly"),
 * the same answer an incomplete plaintext request gets. */
static uint32_t tls_sni_verdict(const orc_tables *t, const uint8_t *p, uint32_t n)
{
    if (n < 5 + 4 || p[0] != 0x16 || p[1] != 0x03) return CLK_N_DENY;
    uint32_t rlen = be16(p + 3);
    if (5 + rlen > n) return CLK_N_NA;                            /* record continues in a later segment */
    const uint8_t *h = p + 5; uint32_t hn = rlen;
    if (h[0] != 1) return CLK_N_DENY;
    uint32_t hl = ((uint32_t)h[1] << 16) | ((uint32_t)h[2] << 8) | h[3];
    if (4 + hl > hn) return CLK_N_NA;                             /* handshake message continues in the next record */
    const uint8_t *b = h + 4; uint32_t o = 0;
    if (hl < 2 + 32 + 1) return CLK_N_DENY;
    o = 34;
    uint32_t sl = b[o]; o += 1 + sl; if (o + 2 > hl) return CLK_N_DENY;
    uint32_t cl = be16(b + o); o += 2 + cl; if (o + 1 > hl) return CLK_N_DENY;
    uint32_t ml = b[o]; o += 1 + ml; if (o + 2 > hl) return CLK_N_DENY;
    uint32_t el = be16(b + o); o += 2; if (o + el != hl) return CLK_N_DENY;
    const uint8_t *sni = NULL; uint32_t snl = 0; int seen = 0;
    while (o + 4 <= hl) {
        uint32_t et = be16(b + o), xl = be16(b + o + 2); o += 4;
        if (o + xl > hl) return CLK_N_DENY;
        if (et == 0) {
            if (seen) return CLK_N_DENY;                          /* duplicate extension */
            seen = 1;
            if (xl < 5) return CLK_N_DENY;
            uint32_t ll = be16(b + o);
            if (ll + 2 != xl || b[o + 2] != 0) return CLK_N_DENY;
            snl = be16(b + o + 3);
            if (snl == 0 || snl > 255 || snl + 3 != ll) return CLK_N_DENY;
            sni = b + o + 5;
        }
        o += xl;
    }
    if (o != hl || !sni) return CLK_N_DENY;
    uint8_t low[256];
    for (uint32_t i = 0; i < snl; i++) { uint8_t c = sni[i]; if (c == 0) return CLK_N_DENY; low[i] = (c >= 'A' && c <= 'Z') ? c + 32 : c; }
    return sni_match(t, low, snl) ? CLK_N_ALLOW : CLK_N_DENY;
}


/* ---- plaintext HTTP on the egress listener: Host header -> virtual host -> first matching path prefix ---------
 * (tables: envoy.go:403-470,566-607,932-944.)  The matcher is Envoy's (v1.37, third party, not in
 * the reference tree: PARITY UNPINNED): RouteMatcher::findVirtualHost looks the lower-cased host up as an exact
 * domain, then against suffix wildcards "*X" (host ends with X, longest X first, the '*' never matches the empty
 * string), then prefix wildcards "X*" (longest first), then "*" — here the implicit deny_all host; inside a
 * virtual host the first route whose prefix matches :path decides (prefix match is case sensitive).  What is
 * stated here for one TCP segment: the request line and the header block up to Host must be in this segment;
 *   - no CRLF yet, or the header block ends with the segment before a Host line was seen  -> N/A (needs reassembly)
 *   - not an HTTP/1.x request line (codec error 400), CONNECT, or no virtual host / route  -> DENY
 *   - absolute-form target "http://authority/path": the authority replaces Host (HTTP/1.1 codec behaviour). */
static int vhost_find(const orc_tables *t, const uint8_t *host, uint32_t hn)
{
    for (uint32_t v = 0; v < t->n_vhosts; v++)                                   /* 1. exact */
        for (uint32_t k = 0; k < t->vhosts[v].n_domains; k++) {
            const char *d = t->vhosts[v].domains[k]; size_t dn = strlen(d);
            if (dn && d[0] != '*' && d[dn - 1] != '*' && dn == hn && memcmp(d, host, hn) == 0) return (int)v;
        }
    int best = -1; size_t best_len = 0;
    for (uint32_t v = 0; v < t->n_vhosts; v++)                                   /* 2. suffix wildcards, longest first */
        for (uint32_t k = 0; k < t->vhosts[v].n_domains; k++) {
            const char *d = t->vhosts[v].domains[k]; size_t dn = strlen(d);
            if (dn < 2 || d[0] != '*') continue;
            size_t sn = dn - 1;
            if (hn > sn && memcmp(host + hn - sn, d + 1, sn) == 0 && sn > best_len) { best = (int)v; best_len = sn; }
        }
    if (best >= 0) return best;
    for (uint32_t v = 0; v < t->n_vhosts; v++)                                   /* 3. prefix wildcards, longest first */
        for (uint32_t k = 0; k < t->vhosts[v].n_domains; k++) {
            const char *d = t->vhosts[v].domains[k]; size_t dn = strlen(d);
            if (dn < 2 || d[0] == '*' || d[dn - 1] != '*') continue;
            size_t pn = dn - 1;
            if (hn > pn && memcmp(host, d, pn) == 0 && pn > best_len) { best = (int)v; best_len = pn; }
        }
    return best;                                                                 /* -1: only "*" (deny_all) is left */
}

static int http_token_char(uint8_t c) { return c > 0x20 && c < 0x7f && !strchr("()<>@,;:\\\"/[]?={}", c); }

static uint32_t http_host_verdict(const orc_tables *t, const uint8_t *p, uint32_t n)
{
    uint32_t eol = 0;
    while (eol + 1 < n && !(p[eol] == '\r' && p[eol + 1] == '\n')) eol++;
    if (eol + 1 >= n) return CLK_N_NA;                                            /* request line not complete */
    uint32_t m = 0;
    while (m < eol && http_token_char(p[m])) m++;
    if (m == 0 || m >= eol || p[m] != ' ') return CLK_N_DENY;
    uint32_t ts = m + 1, te = ts;
    while (te < eol && p[te] != ' ') te++;
    if (te == ts || te >= eol) return CLK_N_DENY;
    if (eol - te != 9 || memcmp(p + te + 1, "HTTP/1.", 7) != 0 || (p[te + 8] != '0' && p[te + 8] != '1')) return CLK_N_DENY;
    if (m == 7 && memcmp(p, "CONNECT", 7) == 0) return CLK_N_DENY;
    const uint8_t *host = NULL; uint32_t hn = 0; const uint8_t *path = p + ts; uint32_t pn = te - ts;
    if (pn >= 7 && (p[ts] | 32) == 'h' && (p[ts + 1] | 32) == 't' && (p[ts + 2] | 32) == 't' && (p[ts + 3] | 32) == 'p' && memcmp(p + ts + 4, "://", 3) == 0) {
        uint32_t a = ts + 7, b = a;
        while (b < te && p[b] != '/') b++;
        host = p + a; hn = b - a;
        if (b < te) { path = p + b; pn = te - b; } else { path = (const uint8_t *)"/"; pn = 1; }
    }
    int complete = 0;
    for (uint32_t ls = eol + 2; host == NULL;) {                                  /* header lines */
        uint32_t le = ls;
        while (le + 1 < n && !(p[le] == '\r' && p[le + 1] == '\n')) le++;
        if (le + 1 >= n) break;                                                   /* line not complete in this segment */
        if (le == ls) { complete = 1; break; }                                    /* end of the header block */
        if (le - ls >= 5 && (p[ls] | 32) == 'h' && (p[ls + 1] | 32) == 'o' && (p[ls + 2] | 32) == 's' && (p[ls + 3] | 32) == 't' && p[ls + 4] == ':') {
            uint32_t a = ls + 5, b = le;
            while (a < b && (p[a] == ' ' || p[a] == '\t')) a++;
            while (b > a && (p[b - 1] == ' ' || p[b - 1] == '\t')) b--;
            host = p + a; hn = b - a;
            break;
        }
        ls = le + 2;
    }
    if (host == NULL) { if (!complete) return CLK_N_NA; hn = 0; host = p; }       /* no Host at all: only "*" can match */
    if (hn > 255) return CLK_N_DENY;
    uint8_t low[256];
    for (uint32_t i = 0; i < hn; i++) low[i] = (host[i] >= 'A' && host[i] <= 'Z') ? host[i] + 32 : host[i];
    int v = hn ? vhost_find(t, low, hn) : -1;
    if (v < 0) return CLK_N_DENY;                                                 /* deny_all: 403 */
    const http_vhost *vh = &t->vhosts[v];
    for (uint32_t r = 0; r < vh->n_routes; r++) {
        size_t rn = strlen(vh->prefixes[r]);
        if (pn >= rn && memcmp(path, vh->prefixes[r], rn) == 0) return vh->allow[r] ? CLK_N_ALLOW : CLK_N_DENY;
    }
    return CLK_N_DENY;                                                            /* no route: 404 */
}

/* tls_inspector: a segment that does not start a TLS handshake record is "raw_buffer" and goes to the HTTP chain
 * when one is configured (envoy.go:358-372); without HTTP rules no filter chain matches and the connection is reset. */
static uint32_t egress_payload_verdict(const orc_tables *t, const uint8_t *p, uint32_t n)
{
    if (t->n_vhosts && n >= 1 && p[0] != 0x16) return http_host_verdict(t, p, n);
    return tls_sni_verdict(t, p, n);
}

/* ------------------------------------------------------------------------------------------------
 * packet -> event mapping and verdict word
 * ---------------------------------------------------------------------------------------------- */

#define ORC_F_CIDR_LINEAR 0x100u      /* oracle-only flag: evaluate A12 by literal linear scan */

static uint32_t target_of(const clk_container_config *cfg, uint32_t ip, uint16_t port)
{
    if (!cfg) return CLK_T_NONE;
    if (ip == cfg->coredns_ip && port == 53) return CLK_T_COREDNS;
    if (ip == cfg->envoy_ip && port == cfg->egress_port) return CLK_T_EGRESS;
    if (ip == cfg->envoy_ip) return CLK_T_TCP;
    return CLK_T_NONE;
}

static uint32_t finish_word(const orc_tables *t, const clk_event *ev, const decision *d, uint32_t name, uint32_t nupd)
{
    uint32_t w = 0;
    w |= d->verdict; w |= (uint32_t)d->metric << 2; w |= name << 4;
    if (d->verdict == CLK_V_REWRITE) {
        const clk_container_config *cfg = (const clk_container_config *)hmap_get(&t->container, &ev->cgroup_id);
        w |= target_of(cfg, d->new_ip, d->new_port) << 6;
        w |= (uint32_t)d->new_port << 8;
    }
    w |= (uint32_t)ev->hook << 24; w |= nupd << 27;
    return w;
}

/* Event-level entry: clawker.c:69-240 on one pre-parsed event. */
uint32_t orc_classify_event(const orc_tables *t, orc_metrics *mx, const clk_event *ev, uint32_t flags)
{
    decision d; memset(&d, 0, sizeof d);
    run_hooks(t, (flags & CLK_F_NO_METRICS) ? NULL : mx, ev, !!(flags & ORC_F_CIDR_LINEAR), &d);
    uint32_t ip_in = (ev->hook >= CLK_HOOK_CONNECT6 && ev->hook <= CLK_HOOK_RECVMSG6) ? ev->user_ip6[3] : ev->user_ip4;
    if (d.verdict == CLK_V_REWRITE && d.new_ip == ip_in && d.new_port == ev->user_port) d.verdict = CLK_V_PASS;  /* identity */
    return finish_word(t, ev, &d, 0, 0);
}

uint32_t orc_classify_packet(const orc_tables *t, orc_metrics *mx, const uint8_t *pkt, uint32_t avail, uint32_t flags,
                             uint32_t pkt_index, clk_dns_update_rec *upd_out)
{
    clk_event ev; memset(&ev, 0, sizeof ev);
    if (avail < 14) return 0;
    uint32_t et = be16(pkt + 12), l4off, proto, end, sport = 0, dport = 0;
    const clk_endpoint *ep = NULL; int v6 = 0, ingress = 0;
    if (et == 0x0800) {
        if (avail < 34 || (pkt[14] >> 4) != 4) return 0;
        uint32_t ihl = pkt[14] & 15; if (ihl < 5) return 0;
        if (be16(pkt + 20) & 0x1fff) return 0;                     /* non-first fragment: no L4 header */
        l4off = 14 + 4 * ihl; proto = pkt[23];
        end = 14 + be16(pkt + 16); if (end > avail) end = avail;
        uint32_t src = ld32(pkt + 26), dst = ld32(pkt + 30);
        if (proto == 6 || proto == 17) {
            if (l4off + 4 > avail) return 0;
            sport = be16(pkt + l4off); dport = be16(pkt + l4off + 2);
        }
        ep = (const clk_endpoint *)hmap_get(&t->endpoint, &src);
        if (!ep && proto == 17) { ep = (const clk_endpoint *)hmap_get(&t->endpoint, &dst); ingress = ep != NULL; }
        ev.user_ip4 = ingress ? src : dst;
        /* DNS answers arriving at CoreDNS from its upstream (no endpoint on either side) */
        if (!ep && proto == 17 && sport == 53 && (flags & CLK_F_NAME_STAGE)) {
            for (uint32_t i = 0; i < t->n_coredns; i++) if (t->coredns_ips[i] == dst) {
                uint32_t po = l4off + 8, nv;
                if (po > end) return 0;
                uint32_t nupd = dns_answer_updates(t, pkt + po, end - po, pkt_index, upd_out, &nv);
                return (nv << 4) | (nupd << 27);
            }
        }
    } else if (et == 0x86DD) {
        if (avail < 54 || (pkt[14] >> 4) != 6) return 0;
        v6 = 1; l4off = 54; proto = pkt[20];
        end = 54 + be16(pkt + 18); if (end > avail) end = avail;
        if (proto == 6 || proto == 17) {
            if (l4off + 4 > avail) return 0;
            sport = be16(pkt + l4off); dport = be16(pkt + l4off + 2);
        }
        ep = (const clk_endpoint *)hmap_get(&t->endpoint6, pkt + 22);
        if (!ep && proto == 17) { ep = (const clk_endpoint *)hmap_get(&t->endpoint6, pkt + 38); ingress = ep != NULL; }
        memcpy(ev.user_ip6, pkt + (ingress ? 22 : 38), 16);
    } else return 0;
    if (!ep) return 0;

    ev.cgroup_id = ep->cgroup_id; ev.uid = ep->uid; ev.so_mark = ep->so_mark;
    if (proto == 6)       { ev.hook = v6 ? CLK_HOOK_CONNECT6 : CLK_HOOK_CONNECT4; ev.sock_type = SOCK_STREAM_; ev.user_port = (uint16_t)dport; }
    else if (proto == 17) {
        ev.sock_type = SOCK_DGRAM_;
        if (ingress) { ev.hook = v6 ? CLK_HOOK_RECVMSG6 : CLK_HOOK_RECVMSG4; ev.user_port = (uint16_t)sport; }
        else         { ev.hook = v6 ? CLK_HOOK_SENDMSG6 : CLK_HOOK_SENDMSG4; ev.user_port = (uint16_t)dport; }
    } else                { ev.hook = CLK_HOOK_SOCK_CREATE; ev.sock_type = SOCK_RAW_; }

    decision d; memset(&d, 0, sizeof d);
    run_hooks(t, (flags & CLK_F_NO_METRICS) ? NULL : mx, &ev, !!(flags & ORC_F_CIDR_LINEAR), &d);
    uint32_t ip_in = v6 ? ev.user_ip6[3] : ev.user_ip4;
    if (d.verdict == CLK_V_REWRITE && d.new_ip == ip_in && d.new_port == ev.user_port) d.verdict = CLK_V_PASS;

    uint32_t name = CLK_N_NA;
    if ((flags & CLK_F_NAME_STAGE) && !ingress && d.verdict != CLK_V_DENY && (proto == 6 || proto == 17)) {
        const clk_container_config *cfg = (const clk_container_config *)hmap_get(&t->container, &ev.cgroup_id);
        int mapped = !v6 || (ev.user_ip6[0] == 0 && ev.user_ip6[1] == 0 && ev.user_ip6[2] == 0xffff0000u);
        if (cfg && mapped) {
            uint32_t fip = d.verdict == CLK_V_REWRITE ? d.new_ip : ip_in;
            uint32_t fport = d.verdict == CLK_V_REWRITE ? d.new_port : ev.user_port;
            uint32_t po = 0, pl = 0;
            if (proto == 17) { po = l4off + 8; if (po <= end) pl = end - po; }
            else if (l4off + 13 <= avail) { uint32_t doff = (pkt[l4off + 12] >> 4) * 4; po = l4off + doff; if (doff >= 20 && po <= end) pl = end - po; }
            if (fip == cfg->coredns_ip && fport == 53) {
                if (proto == 17) name = dns_query_verdict(t, pkt + po, pl);
                else if (pl > 0) name = pl >= 2 && be16(pkt + po) + 2 <= pl ? dns_query_verdict(t, pkt + po + 2, be16(pkt + po)) : CLK_N_DENY;
            } else if (proto == 6 && fip == cfg->envoy_ip && fport == cfg->egress_port && pl > 0) {
                name = egress_payload_verdict(t, pkt + po, pl);
            }
        }
    }
    return finish_word(t, &ev, &d, name, 0);
}

/* ------------------------------------------------------------------------------------------------
 * batch drivers (single thread and pthread-sharded; one metrics map per thread = PERCPU hash)
 * ---------------------------------------------------------------------------------------------- */

typedef struct job {
    const orc_tables *t; orc_metrics *mx;
    const uint8_t *pkts; const uint32_t *offs16; uint32_t stride; uint64_t lo, hi; uint32_t flags;
    uint32_t *verdicts;
} job;

static void *job_run(void *arg)
{
    job *j = (job *)arg;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        const uint8_t *p; uint32_t avail;
        if (j->offs16) { p = j->pkts + (uint64_t)j->offs16[i] * 16; avail = (j->offs16[i + 1] - j->offs16[i]) * 16; }
        else { p = j->pkts + i * j->stride; avail = j->stride; }
        j->verdicts[i] = orc_classify_packet(j->t, j->mx, p, avail, j->flags, (uint32_t)i, NULL);
    }
    return NULL;
}

/* offs16 == NULL -> slot mode with `stride`.  nthreads <= 1 runs inline.  metrics (may be NULL)
 * receives the sum over threads. */
void orc_classify_batch(const orc_tables *t, orc_metrics *metrics, const uint8_t *pkts, const uint32_t *offs16,
                        uint32_t stride, uint64_t n, uint32_t flags, uint32_t *verdicts, int nthreads)
{
    if (nthreads <= 1) {
        job j = { t, metrics, pkts, offs16, stride, 0, n, flags, verdicts };
        job_run(&j); return;
    }
    pthread_t *th = calloc(nthreads, sizeof *th); job *js = calloc(nthreads, sizeof *js);
    for (int k = 0; k < nthreads; k++) {
        js[k] = (job){ t, metrics ? orc_metrics_new() : NULL, pkts, offs16, stride, n * k / nthreads, n * (k + 1) / nthreads, flags, verdicts };
        pthread_create(&th[k], NULL, job_run, &js[k]);
    }
    for (int k = 0; k < nthreads; k++) {
        pthread_join(th[k], NULL);
        if (metrics) { orc_metrics_merge(metrics, js[k].mx); orc_metrics_free(js[k].mx); }
    }
    free(th); free(js);
}

/* The CPU-baseline form: the same sharding, but only the parallel section is timed — the threads (the CPUs the
 * hooks would run on) and their private metrics maps (BPF_MAP_TYPE_PERCPU_HASH: one value per CPU, never merged on
 * the hot path) exist before the clock starts, and the per-thread maps are discarded afterwards, not summed.
 * Returns seconds from "all threads released" to "last thread done". */
typedef struct tjob { job j; pthread_barrier_t *go; } tjob;
static void *tjob_run(void *arg) { tjob *tj = (tjob *)arg; pthread_barrier_wait(tj->go); return job_run(&tj->j); }

double orc_classify_batch_timed(const orc_tables *t, const uint8_t *pkts, const uint32_t *offs16, uint32_t stride, uint64_t n,
                                uint32_t flags, uint32_t *verdicts, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = calloc(nthreads, sizeof *th); tjob *js = calloc(nthreads, sizeof *js);
    pthread_barrier_t go; pthread_barrier_init(&go, NULL, (unsigned)nthreads + 1);
    for (int k = 0; k < nthreads; k++) {
        js[k].j = (job){ t, (flags & CLK_F_NO_METRICS) ? NULL : orc_metrics_new(), pkts, offs16, stride, n * k / nthreads, n * (k + 1) / nthreads, flags, verdicts };
        js[k].go = &go;
        pthread_create(&th[k], NULL, tjob_run, &js[k]);
    }
    struct timespec a, b;
    pthread_barrier_wait(&go);
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    for (int k = 0; k < nthreads; k++) if (js[k].j.mx) orc_metrics_free(js[k].j.mx);
    pthread_barrier_destroy(&go); free(th); free(js);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

/* Stabilised CPU-baseline form (bench.py): like orc_classify_batch_timed, plus
 *   - each thread is pinned to one CPU of the process's affinity set (thread k -> k-th allowed CPU, wrapping),
 *   - each thread works on a PRIVATE copy of its slice that it allocated and first-touched itself (so the pages sit on
 *     its own NUMA node, as the socket buffers of a process running on that CPU would),
 *   - the slice is classified `repeat` times inside the timed section, so a sample can be made seconds long.
 * Verdicts of the last pass land in `verdicts`.  Returns seconds from release to the last thread done. */
typedef struct tjob2 { job j; pthread_barrier_t *go; int cpu, repeat; uint8_t *own; uint32_t *own_offs; } tjob2;
static void *tjob2_run(void *arg)
{
    tjob2 *tj = (tjob2 *)arg; job *j = &tj->j;
    if (tj->cpu >= 0) { cpu_set_t cs; CPU_ZERO(&cs); CPU_SET(tj->cpu, &cs); pthread_setaffinity_np(pthread_self(), sizeof cs, &cs); }
    /* private copy of the slice (first touch on this thread) */
    const uint64_t cnt = j->hi - j->lo;
    job local = *j;
    if (cnt) {
        if (j->offs16) {
            const uint64_t b0 = (uint64_t)j->offs16[j->lo] * 16, b1 = (uint64_t)j->offs16[j->hi] * 16;
            tj->own = malloc(b1 - b0 + 16); tj->own_offs = malloc((cnt + 1) * sizeof(uint32_t));
            memcpy(tj->own, j->pkts + b0, b1 - b0);
            for (uint64_t i = 0; i <= cnt; i++) tj->own_offs[i] = j->offs16[j->lo + i] - j->offs16[j->lo];
            local.pkts = tj->own; local.offs16 = tj->own_offs;
        } else {
            tj->own = malloc(cnt * j->stride);
            memcpy(tj->own, j->pkts + j->lo * j->stride, cnt * j->stride);
            local.pkts = tj->own;
        }
        local.lo = 0; local.hi = cnt; local.verdicts = j->verdicts + j->lo;
    }
    pthread_barrier_wait(tj->go);
    for (int r = 0; r < tj->repeat; r++) job_run(&local);
    return NULL;
}

double orc_classify_batch_timed2(const orc_tables *t, const uint8_t *pkts, const uint32_t *offs16, uint32_t stride, uint64_t n,
                                 uint32_t flags, uint32_t *verdicts, int nthreads, int repeat, int pin)
{
    if (nthreads < 1) nthreads = 1;
    if (repeat < 1) repeat = 1;
    int cpus[4096], ncpu = 0;
    if (pin) { cpu_set_t cs; CPU_ZERO(&cs); if (sched_getaffinity(0, sizeof cs, &cs) == 0) for (int c = 0; c < CPU_SETSIZE && ncpu < 4096; c++) if (CPU_ISSET(c, &cs)) cpus[ncpu++] = c; }
    pthread_t *th = calloc(nthreads, sizeof *th); tjob2 *js = calloc(nthreads, sizeof *js);
    pthread_barrier_t go; pthread_barrier_init(&go, NULL, (unsigned)nthreads + 1);
    for (int k = 0; k < nthreads; k++) {
        js[k].j = (job){ t, (flags & CLK_F_NO_METRICS) ? NULL : orc_metrics_new(), pkts, offs16, stride, n * k / nthreads, n * (k + 1) / nthreads, flags, verdicts };
        js[k].go = &go; js[k].repeat = repeat; js[k].cpu = ncpu ? cpus[k % ncpu] : -1;
        pthread_create(&th[k], NULL, tjob2_run, &js[k]);
    }
    struct timespec a, b;
    pthread_barrier_wait(&go);
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    for (int k = 0; k < nthreads; k++) { if (js[k].j.mx) orc_metrics_free(js[k].j.mx); free(js[k].own); free(js[k].own_offs); }
    pthread_barrier_destroy(&go); free(th); free(js);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

void orc_classify_events(const orc_tables *t, orc_metrics *metrics, const clk_event *ev, uint64_t n, uint32_t flags, uint32_t *verdicts)
{
    for (uint64_t i = 0; i < n; i++) verdicts[i] = orc_classify_event(t, metrics, &ev[i], flags);
}

/* dnsbpf update stream of a varlen batch, packet order then RR order. Returns records written. */
uint64_t orc_dns_answers(const orc_tables *t, const uint8_t *pkts, const uint32_t *offs16, uint64_t n,
                         clk_dns_update_rec *out, uint64_t cap)
{
    uint64_t k = 0; clk_dns_update_rec tmp[CLK_MAX_A_PER_PACKET];
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *p = pkts + (uint64_t)offs16[i] * 16; uint32_t avail = (offs16[i + 1] - offs16[i]) * 16;
        uint32_t w = orc_classify_packet(t, NULL, p, avail, CLK_F_NAME_STAGE | CLK_F_NO_METRICS, (uint32_t)i, tmp);
        uint32_t c = CLK_NUPD(w);
        if (CLK_HOOK(w) != CLK_HOOK_NONE) continue;
        for (uint32_t r = 0; r < c && k < cap; r++) out[k++] = tmp[r];
    }
    return k;
}
