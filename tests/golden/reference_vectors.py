"""Result vectors pinned by the reference's OWN unit tests, transcribed as data (no reference code).

Each block names the reference test it was read from (paths relative to the clawker repository).
These vectors pin the oracle's restatement of the Go rule compiler (oracle/rules_oracle.py) and the
product's C++ rule compiler (clk_rules_compile / clk_validate_dst / clk_domain_hash) alike.
"""

# internal/firewall/rules_test.go:144-218  TestValidateDst  (dst, want_error)
VALIDATE_DST = [
    ("example.com", False), ("api.github.com", False), ("a.b.c.example.com", False), (".example.com", False),
    ("example.com.", False), (".example.com.", False), ("my-api.example.com", False), ("api2.example.com", False),
    ("123.example.com", False), ("localhost", False), ("_dmarc.example.com", False), ("123-456", False),
    ("EXAMPLE.COM", True), ("Api.GitHub.Com", True), (".EXAMPLE.COM", True), ("api.example.co.uk", False),
    ("my.example.technology", False),
    ("a" * 63 + "." + "b" * 63 + "." + "c" * 63 + "." + "d" * 61, False),
    ("a" * 63 + "." + "b" * 63 + "." + "c" * 63 + "." + "d" * 62, True),
    ("192.168.1.1", False), ("::1", False), ("10.0.0.0/8", False), (".192.168.1.1", True), (".10.0.0.0/8", True),
    (".::1", True), ("", True), (".", True), ("..", True), ("example .com", True), ("-example.com", True),
    ("example-.com", True), ("example!.com", True), ("user@example.com", True), ("example.com/path", True),
    ("example.com:443", True), ("https://example.com", True), ("example..com", True), ("example.com..", True),
    (".example.com..", True), ("a" * 64 + ".com", True), ("123.456", True),
]

# internal/firewall/envoy_test.go:15-157  TestTCPMappings  (EnvoyPorts{EgressPort 10000, TCPPortBase 10001})
# (name, rules[(dst, proto, port, action)], expected[(dst, dst_port, envoy_port)])
TCP_MAPPINGS = [
    ("no rules", [], []),
    ("TLS rules are excluded", [("api.anthropic.com", "tls", 443, "allow")], []),
    ("SSH rule", [("github.com", "ssh", 22, "allow")], [("github.com", 22, 10001)]),
    ("TCP rule", [("db.example.com", "tcp", 5432, "allow")], [("db.example.com", 5432, 10001)]),
    ("multiple TCP rules get sequential ports",
     [("github.com", "ssh", 22, "allow"), ("db.example.com", "tcp", 5432, "allow"), ("api.example.com", "tcp", 8080, "allow")],
     [("github.com", 22, 10001), ("db.example.com", 5432, 10002), ("api.example.com", 8080, 10003)]),
    ("mixed TLS and TCP",
     [("api.anthropic.com", "tls", 443, "allow"), ("github.com", "ssh", 22, "allow"), ("registry.npmjs.org", "tls", 443, "allow"),
      ("db.example.com", "tcp", 5432, "allow")],
     [("github.com", 22, 10001), ("db.example.com", 5432, 10002)]),
    ("deny rules are excluded", [("evil.com", "tcp", 8080, "deny"), ("github.com", "ssh", 22, "allow")], [("github.com", 22, 10001)]),
    ("IP/CIDR destinations are excluded", [("192.168.1.0/24", "tcp", 8080, "allow"), ("github.com", "ssh", 22, "allow")],
     [("github.com", 22, 10001)]),
    ("SSH with no port defaults to 22", [("github.com", "ssh", 0, "allow")], [("github.com", 22, 10001)]),
    ("wildcard SSH rule normalized", [(".github.com", "ssh", 22, "allow")], [("github.com", 22, 10001)]),
    ("wildcard TCP rule normalized", [(".db.example.com", "tcp", 5432, "allow")], [("db.example.com", 5432, 10001)]),
]

# internal/firewall/coredns_test.go:14-47 + testdata/corefile_basic.golden: zone blocks in file order
COREFILE_RULES = [("github.com", "", 0, "allow"), ("api.github.com", "", 0, "allow"), ("registry.npmjs.org", "", 0, "allow"),
                  ("10.0.0.0/8", "", 0, "allow"), ("192.168.1.1", "", 0, "allow"), ("evil.example.com", "", 0, "deny"),
                  ("proxy.golang.org", "", 0, "allow"), ("github.com", "", 0, "allow")]
COREFILE_ZONES = ["github.com", "api.github.com", "registry.npmjs.org", "proxy.golang.org",
                  "docker.internal", "otel-collector", "jaeger", "prometheus", "loki", "grafana"]
# coredns_test.go:49-63 TestGenerateCorefile_WildcardDomain
COREFILE_WILDCARD_RULES = [(".datadoghq.com", "", 0, "allow"), ("api.anthropic.com", "", 0, "allow")]
COREFILE_WILDCARD_ZONES_PREFIX = ["datadoghq.com", "api.anthropic.com"]

# envoy_test.go:527-570 TestServerNames_* / TestHTTPDomains_*  (dst, exact set, expected)
SERVER_NAMES = [("api.anthropic.com", [], ["api.anthropic.com"]), (".datadoghq.com", [], [".datadoghq.com", "datadoghq.com"]),
                (".claude.ai", ["claude.ai"], [".claude.ai"])]
HTTP_DOMAINS = [("api.anthropic.com", [], ["api.anthropic.com", "api.anthropic.com:*"]),
                (".datadoghq.com", [], ["*.datadoghq.com", "*.datadoghq.com:*", "datadoghq.com", "datadoghq.com:*"]),
                (".claude.ai", ["claude.ai"], ["*.claude.ai", "*.claude.ai:*"])]

# envoy_test.go:460-525,1001-1014 TestNormalizeAndDedup*  (rules, expected [(dst, port)], warnings)
NORMALIZE_DEDUP = [
    ([("api.anthropic.com", "tls", 0, "allow"), ("github.com", "tls", 0, "allow"), ("api.anthropic.com", "tls", 443, "allow"),
      ("github.com", "tls", 443, "allow")], [("api.anthropic.com", 443), ("github.com", 443)], 0),
    ([("claude.ai", "tls", 443, "allow"), (".claude.ai", "tls", 443, "allow")], [("claude.ai", 443), (".claude.ai", 443)], 0),
    ([("claude.ai", "tls", 443, "allow"), ("claude.ai", "tls", 443, "allow")], [("claude.ai", 443)], 0),
    ([(".", "tls", 443, "allow"), ("..", "tls", 443, "allow"), ("valid.com", "tls", 443, "allow")], [("valid.com", 443)], 2),
]
# envoy_test.go:507-525
IS_WILDCARD = [(".datadoghq.com", True), (".example.com", True), ("api.anthropic.com", False), ("datadoghq.com", False), ("", False)]
NORMALIZE_DOMAIN = [(".datadoghq.com", "datadoghq.com"), (".example.com.", "example.com"), ("api.anthropic.com", "api.anthropic.com"),
                    ("api.anthropic.com.", "api.anthropic.com")]

# internal/ebpf/manager_test.go:67-97 TestDomainHash_CaseInsensitive (relations) ...
DOMAIN_HASH_EQUAL = [(["github.com", "GitHub.com", "GITHUB.COM", "github.COM"], "github.com"), ([".Example.COM", ".example.com"], ".example.com")]
# ... and absolute values from the FNV-1a-32 definition (Go hash/fnv; standard test vectors "", "a", "foobar")
DOMAIN_HASH_KAT = [("", 0x811C9DC5), ("a", 0xE40C292C), ("foobar", 0xBF9CF968), ("github.com", 0x49361819),
                   ("example.com", 0x431CEB26), (".example.com", 0x12445C02)]
# envoy_test.go:159-181 TestTCPMappings_WildcardDomainHashMatchesDnsbpfZone
WILDCARD_ROUTE_RULE = (".example.com", "ssh", 22, "allow")

# internal/dnsbpf/dnsbpf_test.go:204-217 TestZoneToDomain
ZONE_TO_DOMAIN = [("github.com.", "github.com"), (".example.com.", ".example.com"), (".", ""), ("api.github.com.", "api.github.com")]
# dnsbpf_test.go:62-202: (zone, answers[(name, type, ttl, rdata)], expected Update calls [(ip dotted, hash-of, ttl)])
SERVE_DNS = [
    ("github.com.", [("github.com.", "A", 300, "140.82.121.4")], [("140.82.121.4", "github.com", 300)]),
    ("github.com.", [("github.com.", "A", 60, "140.82.121.3"), ("github.com.", "A", 60, "140.82.121.4"), ("github.com.", "AAAA", 60, "::1")],
     [("140.82.121.3", "github.com", 60), ("140.82.121.4", "github.com", 60)]),
    (".example.com.", [("sub.example.com.", "A", 120, "93.184.216.34")], [("93.184.216.34", ".example.com", 120)]),
    ("example.com.", [("example.com.", "A", 0, "93.184.216.34")], [("93.184.216.34", "example.com", 60)]),
]
# internal/ebpf/types.go:86-92 worked example in SURVEY A8 (little-endian host)
IP_TO_U32 = [("140.82.121.4", 75059852), ("127.0.0.1", 0x0100007F), ("not-an-ip", 0)]
