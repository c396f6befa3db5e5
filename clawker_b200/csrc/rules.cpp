// rules.cpp — the rule compiler of clawker's firewall, restated in C++ for the GPU engine.
//
// What the reference does with []config.EgressRule before any traffic flows
// (internal/firewall/manager.go:753-836 ensureConfigs, :1356-1377 syncRoutes):
//   normalizeAndDedup (rules.go:138-161)
//     -> GenerateCorefile zones          (coredns.go:26-96)    => DNS allow zones
//     -> GenerateEnvoyConfig server_names (envoy.go:168-209,358-393,905-915) => SNI allow table
//     -> TCPMappings + DomainHash         (envoy.go:54-83, types.go:121-125) => route_map rows
// plus the A12 extension: IP/CIDR allow rules, which the reference validates and then ignores
// (SURVEY.md R4), become the CIDR allow set.
#include "rules.h"

#include <arpa/inet.h>
#include <string.h>
#include <algorithm>
#include <set>
#include <unordered_set>

namespace clk {

uint32_t domain_hash(const char *s, size_t n)
{
    uint32_t h = 0x811c9dc5u;                                  // hash/fnv New32a offset basis
    for (size_t i = 0; i < n; i++) {
        uint8_t c = (uint8_t)s[i];
        if (c >= 'A' && c <= 'Z') c += 32;                     // strings.ToLower (ASCII)
        h ^= c; h *= 0x01000193u;
    }
    return h;
}

static std::string to_lower(std::string s)
{
    for (auto &c : s) if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
    return s;
}

// net.ParseIP for dotted IPv4: exactly four decimal fields 0..255, no leading zeros, no empty fields.
static bool parse_ipv4(const std::string &s, uint8_t out[4])
{
    size_t i = 0;
    for (int f = 0; f < 4; f++) {
        if (i >= s.size() || s[i] < '0' || s[i] > '9') return false;
        size_t start = i; unsigned v = 0;
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') { v = v * 10 + (unsigned)(s[i] - '0'); if (v > 255) return false; i++; }
        if (i - start > 1 && s[start] == '0') return false;
        out[f] = (uint8_t)v;
        if (f < 3) { if (i >= s.size() || s[i] != '.') return false; i++; }
    }
    return i == s.size();
}

// net.ParseIP for IPv6 text (no zone).  glibc inet_pton implements the same RFC 4291 grammar.
static bool parse_ipv6(const std::string &s, uint8_t out[16])
{
    if (s.find(':') == std::string::npos || s.find('%') != std::string::npos) return false;
    return inet_pton(AF_INET6, s.c_str(), out) == 1;
}

static bool parse_ip_any(const std::string &s, int *version, uint8_t out[16])
{
    if (parse_ipv4(s, out)) { *version = 4; return true; }
    if (parse_ipv6(s, out)) { *version = 6; return true; }
    return false;
}

// net.ParseCIDR: "<ip>/<decimal prefix>"; the returned network is the masked address.
static bool parse_cidr(const std::string &s, int *version, uint8_t addr[16], int *plen)
{
    size_t slash = s.find('/');
    if (slash == std::string::npos) return false;
    if (!parse_ip_any(s.substr(0, slash), version, addr)) return false;
    std::string p = s.substr(slash + 1);
    if (p.empty()) return false;
    long v = 0;
    for (char c : p) { if (c < '0' || c > '9') return false; v = v * 10 + (c - '0'); if (v > 0xFFFFFF) return false; }
    if (v > (*version == 4 ? 32 : 128)) return false;
    *plen = (int)v;
    return true;
}

bool ip_to_u32(const std::string &s, uint32_t *out)
{
    uint8_t b[16]; int ver;
    if (!parse_ip_any(s, &ver, b)) { *out = 0; return false; }
    if (ver == 6) {                                            // To4(): only ::ffff:a.b.c.d converts
        static const uint8_t pfx[12] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xff, 0xff };
        if (memcmp(b, pfx, 12) != 0) { *out = 0; return false; }
        memmove(b, b + 12, 4);
    }
    *out = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    return true;
}

static uint32_t mask_nbo(int plen)
{
    uint32_t host = plen == 0 ? 0 : (0xFFFFFFFFu << (32 - plen));
    return __builtin_bswap32(host);
}

bool cidr_to_addr_mask(const std::string &s, uint32_t *addr, uint32_t *mask)
{
    uint8_t b[16]; int ver, plen;
    if (!parse_cidr(s, &ver, b, &plen) || ver != 4) return false;
    uint32_t a = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    *mask = mask_nbo(plen);
    *addr = a & *mask;
    return true;
}

bool is_ip_or_cidr(const std::string &s)
{
    uint8_t b[16]; int ver, plen;
    return parse_ip_any(s, &ver, b) || parse_cidr(s, &ver, b, &plen);
}

std::string normalize_domain(const std::string &d)
{
    size_t a = 0, e = d.size();
    if (e > a && d[a] == '.') a++;                             // strings.TrimPrefix(d, ".")
    if (e > a && d[e - 1] == '.') e--;                         // strings.TrimSuffix(d, ".")
    return d.substr(a, e - a);
}

bool is_wildcard_domain(const std::string &d) { return !d.empty() && d[0] == '.'; }

bool validate_dst(const std::string &dst)
{
    if (dst.empty()) return false;
    std::string n = normalize_domain(dst);
    if (n.empty()) return false;
    if (is_ip_or_cidr(n)) return dst[0] != '.';                // rules.go:45-55
    if (n.size() > 253) return false;
    char last = '.'; bool non_numeric = false; int partlen = 0;
    for (char c : n) {
        if ((c >= 'a' && c <= 'z') || c == '_') { non_numeric = true; partlen++; }
        else if (c >= '0' && c <= '9') { partlen++; }
        else if (c == '-') { if (last == '.') return false; non_numeric = true; partlen++; }
        else if (c == '.') { if (last == '.' || last == '-') return false; if (partlen > 63) return false; partlen = 0; }
        else return false;                                     // uppercase and everything else
        last = c;
    }
    if (last == '.' || last == '-' || partlen > 63) return false;
    return non_numeric;
}

static bool is_allow(const EgressRule &r)
{
    std::string a = to_lower(r.action);
    return a == "allow" || a.empty();
}

static EgressRule normalize_rule(EgressRule r)                 // rules.go:113-125
{
    if (r.proto.empty()) r.proto = "tls";
    if (r.action.empty()) r.action = "allow";
    if (r.port == 0 && to_lower(r.proto) == "tls") r.port = 443;
    return r;
}

static int tcp_default_port(const EgressRule &r)               // envoy.go:671-679
{
    if (r.port != 0) return r.port;
    return to_lower(r.proto) == "ssh" ? 22 : 443;
}

CompiledRules compile_rules(const std::vector<EgressRule> &in, int tcp_port_base)
{
    CompiledRules out;
    // normalizeAndDedup, rules.go:138-161 (key = dst:proto:port with dst verbatim)
    std::vector<EgressRule> rules;
    {
        std::unordered_set<std::string> seen;
        for (const auto &r0 : in) {
            EgressRule r = normalize_rule(r0);
            if (normalize_domain(r.dst).empty()) { out.warnings++; continue; }
            std::string key = r.dst + ":" + r.proto + ":" + std::to_string(r.port);
            if (!seen.insert(key).second) continue;
            rules.push_back(std::move(r));
        }
    }
    // GenerateCorefile zone set, coredns.go:35-63 + :76-83
    static const char *internal_hosts[] = { "docker.internal", "otel-collector", "jaeger", "prometheus", "loki", "grafana" };
    {
        std::unordered_set<std::string> reserved(std::begin(internal_hosts), std::end(internal_hosts)), emitted;
        for (const auto &r : rules) {
            if (!is_allow(r) || is_ip_or_cidr(r.dst)) continue;            // isAllowDomain coredns.go:98-106
            std::string d = normalize_domain(r.dst);
            if (reserved.count(d) || !emitted.insert(d).second) continue;
            out.zones.push_back(d);
        }
        for (const char *h : internal_hosts) out.zones.push_back(h);
    }
    // GenerateEnvoyConfig classification, envoy.go:176-209; server_names, :905-915
    {
        std::vector<const EgressRule *> tls;
        for (const auto &r : rules) {
            if (!is_allow(r) || is_ip_or_cidr(r.dst)) continue;
            std::string p = to_lower(r.proto);
            if (p == "ssh" || p == "tcp" || p == "http") continue;
            tls.push_back(&r);
        }
        std::unordered_set<std::string> exact, have;
        for (auto r : tls) if (!is_wildcard_domain(r->dst)) exact.insert(normalize_domain(r->dst));      // envoy.go:215-220
        for (auto r : tls) {
            std::string d = normalize_domain(r->dst);
            std::vector<std::string> names;
            if (is_wildcard_domain(r->dst)) { names.push_back("." + d); if (!exact.count(d)) names.push_back(d); }
            else names.push_back(d);
            for (auto &s : names) if (have.insert(s).second) out.snis.push_back(s);
        }
    }
    // raw_buffer (plaintext HTTP) chain: httpRules envoy.go:197-203, httpExactDomains :222-227,
    // httpDomains :932-944, routes :411-425 + buildHTTPRoutes :566-607
    {
        std::vector<const EgressRule *> hs;
        for (const auto &r : rules) if (is_allow(r) && !is_ip_or_cidr(r.dst) && to_lower(r.proto) == "http") hs.push_back(&r);
        std::unordered_set<std::string> exact;
        for (auto r : hs) if (!is_wildcard_domain(r->dst)) exact.insert(normalize_domain(r->dst));
        for (auto r : hs) {
            HttpVhost v; const std::string d = normalize_domain(r->dst);
            if (is_wildcard_domain(r->dst)) {
                v.domains = { "*." + d, "*." + d + ":*" };
                if (!exact.count(d)) { v.domains.push_back(d); v.domains.push_back(d + ":*"); }
            } else v.domains = { d, d + ":*" };
            if (r->path_rules.empty()) v.routes.push_back({ "/", true });
            else {
                for (auto &pr : r->path_rules) v.routes.push_back({ pr.first, to_lower(pr.second) == "allow" });
                const std::string pd = to_lower(r->path_default);
                v.routes.push_back({ "/", !(pd.empty() || pd == "deny") });
            }
            out.http.push_back(std::move(v));
        }
    }
    // TCPMappings, envoy.go:54-83
    {
        int idx = 0;
        for (const auto &r : rules) {
            if (!is_allow(r) || is_ip_or_cidr(r.dst)) continue;
            std::string p = to_lower(r.proto);
            if (p != "ssh" && p != "tcp") continue;
            out.mappings.push_back({ normalize_domain(r.dst), tcp_default_port(r), tcp_port_base + idx });
            idx++;
        }
    }
    // A12 (extension): allow rules whose dst is an IPv4 address or CIDR
    for (const auto &r : rules) {
        if (!is_allow(r)) continue;
        uint8_t b[16]; int ver, plen;
        if (parse_ip_any(r.dst, &ver, b)) {
            if (ver == 4) out.cidrs.push_back({ (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24), 0xFFFFFFFFu });
            continue;
        }
        uint32_t a, m;
        if (parse_cidr(r.dst, &ver, b, &plen) && ver == 4 && cidr_to_addr_mask(r.dst, &a, &m)) out.cidrs.push_back({ a, m });
    }
    return out;
}

}  // namespace clk
