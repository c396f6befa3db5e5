"""Hand-built name-stage packets: (description, frame bytes, expected name verdict or None, expected nupd or None).
Expectations follow DESIGN.md §3.3; the same list is checked against the oracle on CPU and the GPU on B200."""
import struct

from clawker_b200 import capture
from clawker_b200.capture import dns_answer, dns_name_wire, dns_query, frame, ip4, tls_client_hello

CT = ip4("172.20.0.10")          # container #0 of base_tables(): enforced, uid 1000
COREDNS = ip4(capture.COREDNS_IP)
UP = ip4(capture.UPSTREAM_DNS)

RULES = [
    {"dst": "github.com", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": ".example.org", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "claude.ai", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": ".claude.ai", "proto": "tls", "port": 443, "action": "allow"},
    {"dst": "ssh.dev.example.net", "proto": "ssh", "port": 22, "action": "allow"},
    {"dst": "blocked.example.com", "proto": "tls", "port": 443, "action": "deny"},
]


def q(name, **kw):
    return frame(CT, ip4("127.0.0.11"), 17, 40000, 53, dns_query(name, **kw))


def raw_q(payload):
    return frame(CT, ip4("127.0.0.11"), 17, 40000, 53, payload)


def hello(sni, **kw):
    return frame(CT, ip4("93.184.216.34"), 6, 40001, 443, tls_client_hello(sni, **kw))


def ans(name, ips, ttls, **kw):
    return frame(UP, COREDNS, 17, 53, 33333, dns_answer(name, ips, ttls, **kw))


def _hdr(qd=1, an=0, ns=0, ar=0, flags=0x0100):
    return struct.pack(">HHHHHH", 7, flags, qd, an, ns, ar)


ALLOW, DENY, NA = 1, 2, 0
LONG_LABEL = "a" * 64
CASES = [
    # ---- DNS queries: CoreDNS zone selection ----------------------------------------------------------------
    ("exact zone", q("github.com"), ALLOW, None),
    ("subdomain of an exact rule still resolves (zone suffix match)", q("api.github.com"), ALLOW, None),
    ("mixed case", q("GitHub.COM"), ALLOW, None),
    ("wildcard rule: zone is the bare domain", q("a.b.example.org"), ALLOW, None),
    ("wildcard apex", q("example.org"), ALLOW, None),
    ("unlisted", q("evil.example.com"), DENY, None),
    ("deny rule creates no zone", q("blocked.example.com"), DENY, None),
    ("suffix must align on a label", q("notgithub.com"), DENY, None),
    ("tcp/ssh rule still creates a zone", q("ssh.dev.example.net"), ALLOW, None),
    ("internal host zone", q("host.docker.internal"), ALLOW, None),
    ("root name", raw_q(_hdr() + b"\0" + struct.pack(">HH", 1, 1)), DENY, None),
    ("AAAA query", q("github.com", qtype=28), ALLOW, None),
    ("class CH", q("github.com", qclass=3), DENY, None),
    ("response bit set", raw_q(_hdr(flags=0x8100) + dns_name_wire("github.com") + struct.pack(">HH", 1, 1)), DENY, None),
    ("opcode status", raw_q(_hdr(flags=0x1000) + dns_name_wire("github.com") + struct.pack(">HH", 1, 1)), DENY, None),
    ("qdcount 2", raw_q(_hdr(qd=2) + dns_name_wire("github.com") + struct.pack(">HH", 1, 1) * 2), DENY, None),
    ("arcount 3", raw_q(_hdr(ar=3) + dns_name_wire("github.com") + struct.pack(">HH", 1, 1)), DENY, None),
    ("EDNS (arcount 1) accepted", raw_q(_hdr(ar=1) + dns_name_wire("github.com") + struct.pack(">HH", 1, 1) + b"\0\0\x29\x10\0\0\0\0\0\0\0"), ALLOW, None),
    ("truncated header", raw_q(b"\0\x07\x01\0\0\x01"), DENY, None),
    ("truncated name", raw_q(_hdr() + b"\x06githu"), DENY, None),
    ("missing qtype/qclass", raw_q(_hdr() + dns_name_wire("github.com") + b"\0\x01"), DENY, None),
    ("label of 64 octets is malformed", raw_q(_hdr() + bytes([64]) + b"a" * 64 + b"\0" + struct.pack(">HH", 1, 1)), DENY, None),
    ("pointer loop", raw_q(_hdr() + b"\xc0\x0c" + struct.pack(">HH", 1, 1)), DENY, None),
    ("pointer in the question name", raw_q(_hdr() + b"\x03api\xc0\x16" + struct.pack(">HH", 1, 1) + dns_name_wire("github.com")), ALLOW, None),
    ("name of 255+ octets", raw_q(_hdr() + (b"\x3f" + b"a" * 63) * 4 + b"\x03abc\0" + struct.pack(">HH", 1, 1)), DENY, None),
    ("long (>= 64 octets) listed name", q("x" * 40 + "." + "y" * 30 + ".github.com"), ALLOW, None),
    ("label with a dot inside cannot fake a boundary", raw_q(_hdr() + b"\x0aevil.githu" + b"\x01b" + b"\x03com\0" + struct.pack(">HH", 1, 1)), DENY, None),
    ("bad label left of a listed suffix", raw_q(_hdr() + b"\x03a b" + dns_name_wire("github.com") + struct.pack(">HH", 1, 1)), ALLOW, None),
    ("underscore label", q("_dmarc.github.com"), ALLOW, None),
    ("query not to port 53 is not a DNS query", frame(CT, ip4("8.8.8.8"), 17, 40000, 5353, dns_query("github.com")), NA, None),
    ("DNS over TCP with length prefix", frame(CT, ip4("8.8.8.8"), 6, 40000, 53, struct.pack(">H", len(dns_query("api.github.com"))) + dns_query("api.github.com")), ALLOW, None),
    ("DNS over TCP, bad length prefix", frame(CT, ip4("8.8.8.8"), 6, 40000, 53, b"\xff\xff" + dns_query("github.com")), DENY, None),
    ("TCP SYN to port 53 (no payload)", frame(CT, ip4("8.8.8.8"), 6, 40000, 53, b""), NA, None),
    # ---- TLS ClientHello: Envoy server_names ------------------------------------------------------------------
    ("exact SNI", hello("github.com"), ALLOW, None),
    ("exact rule does not cover subdomains", hello("api.github.com"), DENY, None),
    ("wildcard SNI", hello("deep.sub.example.org"), ALLOW, None),
    ("wildcard apex", hello("example.org"), ALLOW, None),
    ("wildcard + exact sibling", hello("x.claude.ai"), ALLOW, None),
    ("exact sibling apex", hello("claude.ai"), ALLOW, None),
    ("upper-case SNI", hello("GITHUB.COM"), ALLOW, None),
    ("unlisted SNI", hello("evil.com"), DENY, None),
    ("no SNI extension", hello(None), DENY, None),
    ("SNI with trailing dot does not match", hello("github.com."), DENY, None),
    ("ssh rule is not a tls rule", hello("ssh.dev.example.net"), DENY, None),
    ("padded hello with session id", hello("github.com", pad_to=480, session_id=bytes(32)), ALLOW, None),
    ("SNI containing NUL", hello("git\0hub.com"), DENY, None),
    ("not TLS (HTTP)", frame(CT, ip4("93.184.216.34"), 6, 40001, 443, b"GET / HTTP/1.1\r\nHost: github.com\r\n\r\n"), DENY, None),
    ("TLS record longer than the segment: needs reassembly", frame(CT, ip4("93.184.216.34"), 6, 40001, 443, tls_client_hello("github.com")[:60]), NA, None),
    ("ClientHello split across two records: needs reassembly", frame(CT, ip4("93.184.216.34"), 6, 40001, 443, (lambda h: h[:3] + (40).to_bytes(2, "big") + h[5:45])(tls_client_hello("github.com"))), NA, None),
    ("SYN to an unlisted address (no payload)", frame(CT, ip4("93.184.216.34"), 6, 40001, 443, b""), NA, None),
    ("TLS to an in-subnet address is not proxied", frame(CT, ip4("172.20.0.77"), 6, 40001, 443, tls_client_hello("evil.com")), NA, None),
    # ---- DNS answers arriving at CoreDNS: dnsbpf ----------------------------------------------------------------
    ("one A record", ans("github.com", [ip4("140.82.121.4")], [300]), ALLOW, 1),
    ("two A + AAAA + CNAME", ans("www.example.org", [ip4("1.2.3.4"), ip4("1.2.3.5")], [0, 60], cname="edge.example.org", aaaa=1), ALLOW, 2),
    ("answer for an unlisted name writes nothing", ans("evil.com", [ip4("6.6.6.6")], [60]), DENY, 0),
    ("A record 0.0.0.0 is skipped", ans("github.com", [0, ip4("140.82.121.4")], [60, 60]), ALLOW, 1),
    ("40 A records: capped at 31", ans("github.com", [ip4("10.9.8.%d" % i) for i in range(1, 41)], [60] * 40), ALLOW, 31),
    ("answer count larger than the message", frame(UP, COREDNS, 17, 53, 33333, dns_answer("github.com", [ip4("1.1.1.1")], [60])[:-3]), ALLOW, 0),
    ("a query (QR=0) from port 53 is not an answer", frame(UP, COREDNS, 17, 53, 33333, dns_query("github.com")), DENY, 0),
    ("answer addressed elsewhere is ignored", frame(UP, ip4("172.20.0.99"), 17, 53, 33333, dns_answer("github.com", [ip4("1.1.1.1")], [60])), NA, 0),
]
