// rules.h — host-side rule compiler: []EgressRule -> zone set / server_names / route rows / CIDR set.
// Mirrors internal/firewall/{rules,coredns,envoy}.go and internal/ebpf/types.go (citations in rules.cpp).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace clk {

struct EgressRule {            // config.EgressRule, internal/config/schema.go:124-138
    std::string dst, proto, action;
    int port = 0;
    std::vector<std::pair<std::string, std::string>> path_rules;   // PathRule {path, action}
    std::string path_default;
};

struct HttpRoute { std::string prefix; bool allow; };                      // buildHTTPRoutes envoy.go:566-607
struct HttpVhost { std::vector<std::string> domains; std::vector<HttpRoute> routes; };   // buildHTTPFilterChain envoy.go:403-425

struct TCPMapping { std::string dst; int dst_port; int envoy_port; };     // firewall.TCPMapping envoy.go:46-50

struct CompiledRules {
    std::vector<std::string> zones;        // Corefile zones in generation order (coredns.go:26-96)
    std::vector<std::string> snis;         // filter_chain_match.server_names in chain order (envoy.go:358-393,905-915)
    std::vector<TCPMapping>  mappings;     // TCPMappings (envoy.go:54-83)
    std::vector<std::pair<uint32_t, uint32_t>> cidrs;   // A12: (addr, mask), NBO-in-native
    std::vector<HttpVhost>   http;         // F3: http_egress_routes virtual hosts in config order (deny_all "*" is implicit)
    uint32_t warnings = 0;                 // rules skipped by normalizeAndDedup (rules.go:146-150)
};

uint32_t domain_hash(const char *s, size_t n);                  // ebpf.DomainHash types.go:121-125
bool     ip_to_u32(const std::string &s, uint32_t *out);        // ebpf.IPToUint32(net.ParseIP(s)) types.go:86-92
bool     cidr_to_addr_mask(const std::string &s, uint32_t *addr, uint32_t *mask);   // types.go:101-111
bool     is_ip_or_cidr(const std::string &s);                   // coredns.go:109-115
std::string normalize_domain(const std::string &d);             // coredns.go:119-122
bool     is_wildcard_domain(const std::string &d);              // coredns.go:126-128
bool     validate_dst(const std::string &dst);                  // rules.go:33-107
CompiledRules compile_rules(const std::vector<EgressRule> &rules, int tcp_port_base);

}  // namespace clk
