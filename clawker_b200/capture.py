"""Synthetic captures and table sets for the BASELINE.json configs.

Everything here is seeded (`numpy.random.Generator(PCG64(seed))`) and vectorised so that the
10 M-packet slot capture of config 2 is produced in seconds.  Two layouts (include/clawker_gpu.h):

  slot mode    [n, stride] uint8, Ethernet II frames zero-padded to `stride` (64)
  varlen mode  one byte blob + offs16[n+1] (16-byte units), each frame 16 B aligned

Table sets are plain dicts ("neutral form") that both the engine front end
(clawker_b200.engine.Engine.load_tables) and the test oracle understand.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

# ---------------------------------------------------------------------------------------------
# address helpers (byte order of internal/ebpf/types.go:86-111)
# ---------------------------------------------------------------------------------------------


def ip4(s: str) -> int:
    a, b, c, d = (int(x) for x in s.split("."))
    return a | (b << 8) | (c << 16) | (d << 24)


def ip4_str(v: int) -> str:
    return ".".join(str((v >> s) & 255) for s in (0, 8, 16, 24))


def prefix_mask(plen: int) -> int:
    """Netmask of a /plen in NBO-in-native-u32 form."""
    if plen == 0:
        return 0
    host = (0xFFFFFFFF << (32 - plen)) & 0xFFFFFFFF
    return int.from_bytes(host.to_bytes(4, "big"), "little")


def fnv1a32_lower(s: str) -> int:
    """ebpf.DomainHash (internal/ebpf/types.go:121-125)."""
    h = 0x811C9DC5
    for c in s.lower().encode():
        h = ((h ^ c) * 0x01000193) & 0xFFFFFFFF
    return h


# ---------------------------------------------------------------------------------------------
# table sets
# ---------------------------------------------------------------------------------------------

NET = "172.20.0.0"
NET_PLEN = 16
ENVOY_IP = "172.20.0.200"      # consts.go: envoy .200, coredns .201
COREDNS_IP = "172.20.0.201"
GATEWAY_IP = "172.20.0.1"
HOST_PROXY_IP = "192.168.65.254"
HOST_PROXY_PORT = 18374
EGRESS_PORT = 10000
TCP_PORT_BASE = 10001
UPSTREAM_DNS = "1.1.1.2"       # coredns.go:12

_WORDS = ("alpha bravo cedar delta ember flint grove harbor indigo juniper kestrel lumen mesa nimbus onyx "
          "pylon quartz river sable tundra umber vertex willow xenon yarrow zephyr anchor beacon cobalt "
          "drift echo fable garnet hollow iris jasper kelp lotus marble nectar opal prism quill raven "
          "summit thistle unity velvet wander yonder zenith").split()
_TLDS = ("com", "net", "org", "io", "dev", "cloud", "app")


def domain_list(n: int, seed: int = 0) -> list[str]:
    """n distinct lower-case domains built from a fixed word list."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    out, seen = [], set()
    W, T = len(_WORDS), len(_TLDS)
    i = 0
    while len(out) < n:
        a, b, t, k = rng.integers(0, W), rng.integers(0, W), rng.integers(0, T), rng.integers(0, 4)
        if k == 0:
            d = f"{_WORDS[a]}{_WORDS[b]}.{_TLDS[t]}"
        elif k == 1:
            d = f"{_WORDS[a]}-{_WORDS[b]}{i % 997}.{_TLDS[t]}"
        elif k == 2:
            d = f"{_WORDS[a]}.{_WORDS[b]}{i % 89}.{_TLDS[t]}"
        else:
            d = f"api{i % 53}.{_WORDS[a]}{_WORDS[b]}.{_TLDS[t]}"
        i += 1
        if d not in seen:
            seen.add(d)
            out.append(d)
    return out


@dataclass
class Workload:
    tables: dict            # neutral table set (without zones/snis/routes/cidrs derived from rules)
    rules: list             # list of dicts dst/proto/port/action for the rule compiler
    meta: dict


def base_tables(n_containers: int = 8) -> dict:
    """Containers 172.20.0.10.. with cgroup ids 1001..; per-container quirks that exercise A1:
    #1 runs as root (uid 0 -> pass), #2 is bypassed, #3 is an endpoint without a container_map row
    (unmanaged), #4 carries CLAWKER_MARK (Envoy-like upstream socket)."""
    cfg = (ip4(ENVOY_IP), ip4(COREDNS_IP), ip4(GATEWAY_IP), ip4(NET), prefix_mask(NET_PLEN),
           ip4(HOST_PROXY_IP), HOST_PROXY_PORT, EGRESS_PORT)
    t = {"containers": {}, "bypass": {}, "dns_cache": {}, "routes": {}, "endpoints": {}, "endpoints6": {},
         "cidrs": [], "zones": [], "snis": []}
    for i in range(n_containers):
        cg = 1001 + i
        ip = ip4(f"172.20.{i // 200}.{10 + i % 200}")
        uid, mark = 1000, 0
        if n_containers >= 6:
            if i == 1:
                uid = 0
            if i == 4:
                mark = 0xC1A4
        t["endpoints"][ip] = (cg, uid, mark)
        ip6 = bytes([0xFD, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, i >> 8, 10 + (i & 255)])
        t["endpoints6"][ip6] = (cg, uid, mark)
        if n_containers >= 6 and i == 3:
            continue
        t["containers"][cg] = cfg
        if n_containers >= 6 and i == 2:
            t["bypass"][cg] = 1
    return t


def make_workload(config: str, seed: int | None = None, n_rules: int | None = None) -> Workload:
    """Table set + rules for a BASELINE.json config: 'c1' (32 allow IPs, 1 container), 'c2' (1k rules:
    800 dns_cache IPs incl. 200 route rows + 200 CIDRs, 8 containers), 'c3' (10k-domain allowlist, 20 %
    wildcard), 'c4' (100k rules), 'c5:<entries>' (rule-table sweep)."""
    name = config.split(":")[0]
    with_http = name == "c3h"                           # c3 + plaintext-HTTP rules for every 8th domain
    if with_http:
        name = "c3"
    seed = {"c1": 1, "c2": 2, "c3": 3, "c4": 4, "c5": 5}[name] if seed is None else seed
    rng = np.random.Generator(np.random.PCG64(seed))
    if name == "c1":
        t = base_tables(1)
        n_ip, n_route, n_cidr, n_dom = 32, 4, 0, 8
    elif name == "c2":
        t = base_tables(8)
        n = n_rules or 1000
        n_ip, n_route, n_cidr, n_dom = n * 8 // 10, n * 2 // 10, n * 2 // 10, 64
    elif name == "c3":
        t = base_tables(8)
        n = n_rules or 10000
        n_ip, n_route, n_cidr, n_dom = min(n, 8192), min(n // 50, 512), 0, n
    elif name == "c4":
        t = base_tables(64)
        n = n_rules or 100000
        n_ip, n_route, n_cidr, n_dom = n // 2, 2000, n // 10, n * 4 // 10
    elif name == "c5":
        t = base_tables(8)
        n = n_rules or int(config.split(":")[1])
        n_ip, n_route, n_cidr, n_dom = n // 2, min(n // 8, 4096), n // 2, 16
    else:
        raise ValueError(config)

    domains = domain_list(n_dom, seed)
    rules = []
    wildcard = rng.random(n_dom) < 0.2
    for i, d in enumerate(domains):
        rules.append({"dst": ("." + d) if wildcard[i] else d, "proto": "tls", "port": 443, "action": "allow"})
    # tcp/ssh rules -> route_map rows (TCPMappings envoy.go:54-83); ports from a small set
    route_ports = [22, 8080, 5432, 6379, 9418]
    route_domains = [domains[int(i)] for i in rng.choice(n_dom, size=min(n_route, n_dom), replace=False)] if n_route else []
    # more routes than domains: add extra ports on the same domains
    r_i = 0
    while len(route_domains) < n_route:
        route_domains.append(domains[r_i % n_dom])
        r_i += 1
    seen_r = set()
    for i, d in enumerate(route_domains):
        port = route_ports[i % len(route_ports)] + (i // (len(route_ports) * max(n_dom, 1)))
        proto = "ssh" if port == 22 else "tcp"
        if (d, port) in seen_r:
            port = 20000 + i
        seen_r.add((d, port))
        rules.append({"dst": d, "proto": proto, "port": port, "action": "allow"})
    if with_http:
        for i, d in enumerate(domains[::8]):
            r = {"dst": ("." + d) if wildcard[i * 8] else d, "proto": "http", "port": 80, "action": "allow"}
            if i % 3 == 0:
                r.update(path_rules=[("/api/", "allow"), ("/admin", "deny")], path_default="deny" if i % 2 else "allow")
            rules.append(r)
    # a few deny rules and duplicates that the compiler must ignore
    rules.append({"dst": "blocked.example.org", "proto": "tls", "port": 443, "action": "deny"})
    if domains:
        rules.append({"dst": domains[0], "proto": "", "port": 0, "action": ""})       # dup after normalisation? (dst verbatim)
    # IP/CIDR allow rules (A12)
    cidr_plens = rng.integers(8, 31, size=n_cidr)
    for p in cidr_plens:
        a = int(rng.integers(0x01000000, 0xDF000000))
        a &= (0xFFFFFFFF << (32 - int(p))) & 0xFFFFFFFF
        o = a.to_bytes(4, "big")
        if o[0] in (127, 172, 10):
            o = bytes([o[0] + 1]) + o[1:]
        rules.append({"dst": f"{o[0]}.{o[1]}.{o[2]}.{o[3]}/{int(p)}", "proto": "tcp", "port": 0, "action": "allow"})

    # dns_cache: resolved IPs -> hash of a listed domain (what dnsbpf would have written)
    ips = rng.integers(0x01000000, 0xDF000000, size=n_ip * 2, dtype=np.int64)
    ips = np.unique(ips)[:n_ip]
    rng.shuffle(ips)
    for i, v in enumerate(ips):
        o = int(v).to_bytes(4, "big")
        if o[0] in (127, 172):
            continue
        ipn = int.from_bytes(o, "little")
        d = route_domains[i % len(route_domains)] if (route_domains and i % 4 == 0) else domains[i % n_dom]
        t["dns_cache"][ipn] = (fnv1a32_lower(d), 0x7FFFFFFF)
    meta = {"config": config, "seed": seed, "n_dns_cache": len(t["dns_cache"]), "n_rules": len(rules),
            "n_domains": n_dom, "n_cidr": n_cidr, "n_routes": n_route, "domains": domains,
            "wildcard": wildcard}
    return Workload(t, rules, meta)


# ---------------------------------------------------------------------------------------------
# packet synthesis
# ---------------------------------------------------------------------------------------------

_MAC_GW = bytes.fromhex("02420a000001")
_MAC_CT = bytes.fromhex("02420a00000a")


def _put16be(a: np.ndarray, col: int, v: np.ndarray) -> None:
    a[:, col] = (v >> 8) & 255
    a[:, col + 1] = v & 255


def _put32ne(a: np.ndarray, col: int, v: np.ndarray) -> None:
    """Write NBO-in-native u32 addresses (little-endian store == network byte order on the wire)."""
    v = v.astype(np.uint32)
    for k in range(4):
        a[:, col + k] = (v >> (8 * k)) & 255


def _dst_pool(w: Workload, rng) -> np.ndarray:
    """Destination IPs drawn 'from the rule set': dns_cache IPs + addresses inside CIDR rules +
    infrastructure addresses (loopback, in-subnet, gateway, host proxy, CoreDNS, docker DNS)."""
    pool = list(w.tables["dns_cache"].keys())
    for r in w.rules:
        d = r["dst"]
        if "/" in d and d[0].isdigit():
            a, p = d.split("/")
            base = int.from_bytes(bytes(int(x) for x in a.split(".")), "big")
            host = int(rng.integers(0, 1 << (32 - int(p)))) if int(p) < 32 else 0
            pool.append(int.from_bytes(((base + host) & 0xFFFFFFFF).to_bytes(4, "big"), "little"))
    infra = [ip4("127.0.0.1"), ip4("127.0.0.11"), ip4("172.20.0.5"), ip4(GATEWAY_IP), ip4(HOST_PROXY_IP),
             ip4(COREDNS_IP), ip4(ENVOY_IP), ip4("8.8.8.8")]
    return np.array(pool + infra * max(1, len(pool) // 64), dtype=np.uint32)


def _l34_fields(w: Workload, n: int, rng, tcp_frac=0.85, udp_frac=0.14, v6_frac=0.02, nonip_frac=0.005):
    """Per-packet L3/L4 field arrays for egress packets of the workload's containers."""
    eps = np.array(list(w.tables["endpoints"].keys()), dtype=np.uint32)
    src = eps[rng.integers(0, eps.size, size=n)]
    # 3 % of packets come from addresses that are not endpoints (unmanaged hosts)
    stranger = rng.random(n) < 0.03
    src = np.where(stranger, rng.integers(0x0A000000, 0x0AFFFFFF, size=n).astype(np.uint32), src)
    pool = _dst_pool(w, rng)
    from_rules = rng.random(n) < 0.5
    dst = np.where(from_rules, pool[rng.integers(0, pool.size, size=n)],
                   rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32))
    # dport: Zipf-like over {443,80,22,53,8080,random} plus the workload's route ports
    route_ports = sorted({r["port"] for r in w.rules if r["proto"] in ("tcp", "ssh") and r["port"]}) or [22]
    choice = rng.choice(7, size=n, p=[0.45, 0.15, 0.1, 0.1, 0.05, 0.05, 0.10])
    fixed = np.array([443, 80, 22, 53, 8080, HOST_PROXY_PORT, 0], dtype=np.uint32)
    dport = fixed[choice]
    rnd = choice == 6
    dport = np.where(rnd, rng.integers(1, 65536, size=n).astype(np.uint32), dport)
    # rule-set destinations often use one of the routed ports
    use_route = from_rules & (rng.random(n) < 0.3)
    dport = np.where(use_route, np.array(route_ports, dtype=np.uint32)[rng.integers(0, len(route_ports), size=n)], dport)
    sport = rng.integers(1024, 65536, size=n).astype(np.uint32)
    u = rng.random(n)
    proto = np.where(u < tcp_frac, 6, np.where(u < tcp_frac + udp_frac, 17, 1)).astype(np.uint8)
    other = proto == 1
    proto = np.where(other & (rng.random(n) < 0.5), 47, proto).astype(np.uint8)      # ICMP / GRE
    fam = rng.random(n)
    v6 = fam < v6_frac
    nonip = (fam >= v6_frac) & (fam < v6_frac + nonip_frac)
    return dict(src=src, dst=dst, sport=sport, dport=dport, proto=proto, v6=v6, nonip=nonip)


def _write_headers(buf: np.ndarray, f: dict, rng, totlen: np.ndarray | None = None) -> None:
    """Fill Ethernet + IPv4/IPv6 + first 20 L4 bytes of each row of buf [n, >=64]."""
    n = buf.shape[0]
    buf[:, 0:6] = np.frombuffer(_MAC_GW, dtype=np.uint8)
    buf[:, 6:12] = np.frombuffer(_MAC_CT, dtype=np.uint8)
    v6, nonip, proto = f["v6"], f["nonip"], f["proto"]
    v4 = ~v6 & ~nonip
    # IPv4
    buf[v4, 12], buf[v4, 13] = 0x08, 0x00
    buf[v4, 14] = 0x45
    if totlen is None:                                                        # header-only frames
        totlen = np.where(proto == 6, 40, 28).astype(np.uint32)
    tl = totlen
    _put16be(buf, 16, np.where(v4, tl, 0).astype(np.uint32))
    buf[v4, 22] = 64
    buf[v4, 23] = proto[v4]
    a = np.zeros((n, 8), dtype=np.uint8)
    _put32ne(a, 0, f["src"])
    _put32ne(a, 4, f["dst"])
    buf[v4, 26:34] = a[v4]
    l4 = np.zeros((n, 20), dtype=np.uint8)
    _put16be(l4, 0, f["sport"])
    _put16be(l4, 2, f["dport"])
    tcp = proto == 6
    l4[tcp, 12] = 0x50
    l4[tcp, 13] = 0x02
    udp = proto == 17
    ul = np.maximum(tl.astype(np.int64) - 20, 8).astype(np.uint32)
    l4u = l4[udp]
    _put16be(l4u, 4, ul[udp])
    l4[udp] = l4u
    l4[~tcp & ~udp] = 0
    l4[~tcp & ~udp, 0] = 8                                                    # ICMP echo
    buf[v4, 34:54] = l4[v4]
    # IPv6: half native (fd00::/8 style), quarter v4-mapped dst, rest ::1
    if v6.any():
        k = int(v6.sum())
        buf[v6, 12], buf[v6, 13] = 0x86, 0xDD
        buf[v6, 14] = 0x60
        b6 = buf[v6]
        _put16be(b6, 18, np.full(k, 8, dtype=np.uint32))
        b6[:, 20] = proto[v6]
        b6[:, 21] = 64
        src6 = np.zeros((k, 16), dtype=np.uint8)
        src6[:, 0] = 0xFD
        s4 = np.zeros((k, 4), dtype=np.uint8)
        _put32ne(s4, 0, f["src"][v6])
        src6[:, 15] = s4[:, 3]                       # last octet of the container's v4 address == base_tables()
        src6[:, 14] = 0
        b6[:, 22:38] = src6
        kind = rng.integers(0, 4, size=k)
        dst6 = rng.integers(0, 256, size=(k, 16)).astype(np.uint8)
        dst6[:, 0] = 0x20
        mapped = kind == 1
        dst6[mapped, 0:10] = 0
        dst6[mapped, 10:12] = 0xFF
        d4 = np.zeros((k, 4), dtype=np.uint8)
        _put32ne(d4, 0, f["dst"][v6])
        dst6[mapped, 12:16] = d4[mapped]
        loop = kind == 2
        dst6[loop] = 0
        dst6[loop, 15] = 1
        b6[:, 38:54] = dst6
        b6[:, 54:58] = l4[v6][:, 0:4]
        buf[v6] = b6
    if nonip.any():
        buf[nonip, 12], buf[nonip, 13] = 0x08, 0x06                           # ARP
    return None


def gen_slots(w: Workload, n: int, seed: int = 0, stride: int = 64, chunk: int = 1 << 20, **mix) -> np.ndarray:
    """n packets in slot mode.  Returns a [n, stride] uint8 array."""
    out = np.zeros((n, stride), dtype=np.uint8)
    rng = np.random.Generator(np.random.PCG64(seed * 1000003 + 17))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        f = _l34_fields(w, hi - lo, rng, **mix)
        _write_headers(out[lo:hi], f, rng)
    return out


def gen_events(w: Workload, n: int, seed: int = 0) -> np.ndarray:
    """Random socket events covering every hook, socket type and address class."""
    from .abi import EVENT_DTYPE
    rng = np.random.Generator(np.random.PCG64(seed * 7 + 3))
    f = _l34_fields(w, n, rng)
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    eps = w.tables["endpoints"]
    keys = np.array(list(eps.keys()), dtype=np.uint32)
    vals = np.array([eps[int(k)] for k in keys], dtype=np.uint64)
    pick = rng.integers(0, keys.size, size=n)
    ev["cgroup_id"] = vals[pick, 0]
    ev["uid"] = vals[pick, 1]
    ev["so_mark"] = vals[pick, 2]
    stranger = rng.random(n) < 0.05
    ev["cgroup_id"] = np.where(stranger, 999999, ev["cgroup_id"])
    ev["hook"] = rng.integers(0, 8, size=n)
    ev["sock_type"] = rng.choice([1, 2, 3, 5], size=n, p=[0.55, 0.35, 0.08, 0.02])
    ev["user_ip4"] = f["dst"]
    ev["user_port"] = f["dport"]
    # recvmsg: source often CoreDNS:53
    rc = (ev["hook"] == 3) | (ev["hook"] == 6)
    from_dns = rc & (rng.random(n) < 0.6)
    ev["user_ip4"] = np.where(from_dns, ip4(COREDNS_IP), ev["user_ip4"])
    ev["user_port"] = np.where(from_dns, 53, ev["user_port"])
    ip6 = rng.integers(0, 1 << 32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    kind = rng.integers(0, 4, size=n)
    mapped = kind <= 1
    ip6[mapped, 0] = 0
    ip6[mapped, 1] = 0
    ip6[mapped, 2] = 0xFFFF0000
    ip6[mapped, 3] = ev["user_ip4"][mapped]
    loop = kind == 2
    ip6[loop] = 0
    ip6[loop, 3] = 0x01000000
    ev["user_ip6"] = ip6
    return ev


# ---- name-stage payloads --------------------------------------------------------------------

def dns_name_wire(name: str) -> bytes:
    out = b""
    for lab in name.split("."):
        if lab:
            out += bytes([len(lab)]) + lab.encode()
    return out + b"\0"


def dns_query(name: str, qtype: int = 1, txid: int = 0x1234, qclass: int = 1) -> bytes:
    return struct.pack(">HHHHHH", txid, 0x0100, 1, 0, 0, 0) + dns_name_wire(name) + struct.pack(">HH", qtype, qclass)


def dns_answer(name: str, ips: list, ttls: list, txid: int = 0x1234, cname: str | None = None,
               aaaa: int = 0) -> bytes:
    """Response with optional CNAME + A RRs (+ AAAA RRs), owner names compressed to the question."""
    an = len(ips) + (1 if cname else 0) + aaaa
    m = struct.pack(">HHHHHH", txid, 0x8180, 1, an, 0, 0) + dns_name_wire(name) + struct.pack(">HH", 1, 1)
    if cname:
        rd = dns_name_wire(cname)
        m += b"\xc0\x0c" + struct.pack(">HHIH", 5, 1, 300, len(rd)) + rd
    for ip, ttl in zip(ips, ttls):
        m += b"\xc0\x0c" + struct.pack(">HHIH", 1, 1, ttl, 4) + int(ip).to_bytes(4, "little")
    for _ in range(aaaa):
        m += b"\xc0\x0c" + struct.pack(">HHIH", 28, 1, 300, 16) + bytes(range(16))
    return m


def tls_client_hello(sni: str | None, pad_to: int = 0, session_id: bytes = b"", extra_ext: bool = True) -> bytes:
    """A TLS 1.2/1.3-style ClientHello record with (optionally) a server_name extension."""
    exts = b""
    if extra_ext:
        exts += struct.pack(">HH", 0x000A, 4) + b"\x00\x02\x00\x1d"            # supported_groups
    if sni is not None:
        host = sni.encode()
        sn = struct.pack(">HBH", len(host) + 3, 0, len(host)) + host
        exts += struct.pack(">HH", 0, len(sn)) + sn
    if extra_ext:
        exts += struct.pack(">HH", 0x0010, 5) + b"\x00\x03\x02h2"              # ALPN
    body = b"\x03\x03" + bytes(range(32)) + bytes([len(session_id)]) + session_id
    body += struct.pack(">H", 4) + b"\x13\x01\x13\x02" + b"\x01\x00"
    cur = 5 + 4 + len(body) + 2 + len(exts)
    if pad_to > cur + 4:
        exts += struct.pack(">HH", 0x0015, pad_to - cur - 4) + bytes(pad_to - cur - 4)   # padding ext
    body += struct.pack(">H", len(exts)) + exts
    hs = b"\x01" + len(body).to_bytes(3, "big") + body
    return b"\x16\x03\x01" + struct.pack(">H", len(hs)) + hs


def _mixed_case(rng, s: str) -> str:
    m = rng.random(len(s)) < 0.3
    return "".join(c.upper() if u else c for c, u in zip(s, m))


def build_varlen(frames: list[bytes]) -> tuple[np.ndarray, np.ndarray]:
    """Pack raw frames into (blob, offs16); every frame padded to >= 64 B and a multiple of 16 B."""
    lens = np.array([max(64, (len(f) + 15) & ~15) for f in frames], dtype=np.int64)
    offs = np.zeros(len(frames) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    blob = np.zeros(int(offs[-1]), dtype=np.uint8)
    for f, o in zip(frames, offs[:-1]):
        blob[o:o + len(f)] = np.frombuffer(f, dtype=np.uint8)
    return blob, (offs // 16).astype(np.uint32)


def frame(src: int, dst: int, proto: int, sport: int, dport: int, payload: bytes = b"", tcp_flags: int = 0x18, seq: int = 1) -> bytes:
    """One Ethernet II / IPv4 / TCP|UDP frame around `payload`."""
    if proto == 6:
        l4 = struct.pack(">HHIIBBHHH", sport, dport, seq & 0xffffffff, 1, 0x50, tcp_flags, 65535, 0, 0)
    elif proto == 17:
        l4 = struct.pack(">HHHH", sport, dport, 8 + len(payload), 0)
    else:
        l4 = b"\x08\x00\x00\x00\x00\x01\x00\x01"
    tot = 20 + len(l4) + len(payload)
    ip = struct.pack(">BBHHHBBH", 0x45, 0, tot, 0, 0x4000, 64, proto, 0) + int(src).to_bytes(4, "little") + int(dst).to_bytes(4, "little")
    return _MAC_GW + _MAC_CT + b"\x08\x00" + ip + l4 + payload


def gen_varlen(w: Workload, n: int, seed: int = 0, pool: int = 4096,
               mix=(0.70, 0.10, 0.05, 0.15)) -> tuple[np.ndarray, np.ndarray, dict]:
    """Config-3 style capture: `mix` = fractions of (64 B TCP/UDP/other headers, DNS queries, DNS answers
    arriving at CoreDNS, TLS ClientHellos[, plaintext HTTP requests]).  Payloads come from seeded pools of `pool` distinct messages;
    per-packet L3/L4 fields are drawn independently.  Returns (blob, offs16, info)."""
    rng = np.random.Generator(np.random.PCG64(seed * 1000003 + 29))
    domains, wildcard = w.meta["domains"], w.meta["wildcard"]
    nd = len(domains)

    def pick_name(kind_hit_exact, kind_hit_sub, miss):
        u = rng.random()
        i = int(rng.integers(0, nd))
        if u < kind_hit_exact:
            return domains[i]
        if u < kind_hit_exact + kind_hit_sub:
            wi = np.flatnonzero(wildcard)
            j = int(wi[rng.integers(0, wi.size)]) if wi.size else i
            return _WORDS[int(rng.integers(0, len(_WORDS)))] + "." + domains[j]
        return _WORDS[int(rng.integers(0, len(_WORDS)))] + str(int(rng.integers(0, 9999))) + ".invalid-" + _TLDS[i % len(_TLDS)]

    q_pool, a_pool, t_pool = [], [], []
    for _ in range(pool):
        nm = pick_name(0.35, 0.15, 0.5)
        extra = int(rng.integers(0, 3))
        for _e in range(extra):
            nm = _WORDS[int(rng.integers(0, len(_WORDS)))] + "." + nm
        q_pool.append(dns_query(_mixed_case(rng, nm), qtype=int(rng.choice([1, 28, 1, 1, 65])), txid=int(rng.integers(0, 65536))))
        nm = pick_name(0.6, 0.25, 0.15)
        k = int(rng.integers(1, 5))
        ips = [int(x) for x in rng.integers(0x01000001, 0xDF000000, size=k)]
        ttls = [int(x) for x in rng.choice([0, 60, 300], size=k)]
        a_pool.append(dns_answer(_mixed_case(rng, nm), ips, ttls, txid=int(rng.integers(0, 65536)),
                                 cname=("edge." + nm) if rng.random() < 0.3 else None, aaaa=int(rng.integers(0, 2))))
        u = rng.random()
        sni = pick_name(0.5, 0.2, 0.3) if u < 0.95 else None
        t_pool.append(tls_client_hello(_mixed_case(rng, sni) if sni else None, pad_to=int(rng.integers(200, 518)) - 54,
                                       session_id=bytes(32) if rng.random() < 0.5 else b""))

    def pool_arrays(p):
        ln = np.array([len(x) for x in p], dtype=np.int64)
        st = np.zeros(len(p) + 1, dtype=np.int64)
        np.cumsum(ln, out=st[1:])
        return np.frombuffer(b"".join(p), dtype=np.uint8), st[:-1], ln

    pools = [pool_arrays(q_pool), pool_arrays(a_pool), pool_arrays(t_pool)]
    if len(mix) == 5:                                   # after the other pools: a 4-class mix draws exactly what it always drew
        h_pool = []
        paths = ["/", "/api/v1/items", "/api/", "/admin/login", "/index.html", "/static/app.js?v=3", "/healthz"]
        for _ in range(pool):
            host = pick_name(0.45, 0.2, 0.35)
            if rng.random() < 0.15:
                host += ":%d" % int(rng.choice([80, 8080, 3000]))
            hdrs = [b"User-Agent: curl/8.5.0", b"Accept: */*"] if rng.random() < 0.7 else [b"Accept-Encoding: gzip", b"Connection: keep-alive", b"X-Request-Id: " + bytes(rng.integers(97, 123, size=24, dtype=np.uint8))]
            host_line = b"Host: " + _mixed_case(rng, host).encode()
            pos = int(rng.integers(0, len(hdrs) + 1))
            lines = hdrs[:pos] + [host_line] + hdrs[pos:]
            h_pool.append(("%s %s HTTP/1.1\r\n" % (str(rng.choice(["GET", "GET", "POST", "HEAD"])), paths[int(rng.integers(0, len(paths)))])).encode()
                          + b"\r\n".join(lines) + b"\r\n\r\n")
        pools.append(pool_arrays(h_pool))
    cls = rng.choice(len(mix), size=n, p=list(mix)).astype(np.int8)
    f = _l34_fields(w, n, rng, v6_frac=0.01, nonip_frac=0.002)
    pidx = rng.integers(0, pool, size=n)
    pay_len = np.zeros(n, dtype=np.int64)
    pay_src = np.zeros(n, dtype=np.int64)
    hdr_len = np.full(n, 54, dtype=np.int64)
    v4 = ~f["v6"] & ~f["nonip"]
    # class 1: DNS query -> UDP dport 53 from a container (10 % over TCP with length prefix omitted: stay UDP)
    c1 = (cls == 1) & v4
    f["proto"] = np.where(c1, 17, f["proto"]).astype(np.uint8)
    f["dport"] = np.where(c1, 53, f["dport"])
    # class 2: DNS answer upstream -> CoreDNS, sport 53
    c2 = (cls == 2) & v4
    f["proto"] = np.where(c2, 17, f["proto"]).astype(np.uint8)
    f["src"] = np.where(c2, ip4(UPSTREAM_DNS), f["src"]).astype(np.uint32)
    f["dst"] = np.where(c2, ip4(COREDNS_IP), f["dst"]).astype(np.uint32)
    f["sport"] = np.where(c2, 53, f["sport"])
    # class 3: TLS ClientHello -> TCP dport 443 mostly
    c3 = (cls == 3) & v4
    f["proto"] = np.where(c3, 6, f["proto"]).astype(np.uint8)
    f["dport"] = np.where(c3 & (rng.random(n) < 0.9), 443, f["dport"])
    classes = [c1, c2, c3]
    if len(mix) == 5:                                   # class 4: plaintext HTTP request -> TCP dport 80
        c4 = (cls == 4) & v4
        f["proto"] = np.where(c4, 6, f["proto"]).astype(np.uint8)
        f["dport"] = np.where(c4, 80, f["dport"])
        classes.append(c4)
    for k, c in enumerate(classes):
        blobk, stk, lnk = pools[k]
        pay_len = np.where(c, lnk[pidx], pay_len)
        pay_src = np.where(c, stk[pidx], pay_src)
    udp = f["proto"] == 17
    hdr_len = np.where(udp & v4, 42, hdr_len)
    totlen = np.where(f["proto"] == 6, 40, 28).astype(np.int64)
    totlen = np.where(pay_len > 0, hdr_len - 14 + pay_len, totlen).astype(np.uint32)
    frame_len = np.maximum(64, (hdr_len + pay_len + 15) & ~15)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(frame_len, out=offs[1:])
    blob = np.zeros(int(offs[-1]), dtype=np.uint8)
    CH = 1 << 18
    for lo in range(0, n, CH):
        hi = min(n, lo + CH)
        hb = np.zeros((hi - lo, 64), dtype=np.uint8)
        sub = {k: v[lo:hi] for k, v in f.items()}
        _write_headers(hb, sub, rng, totlen=totlen[lo:hi])
        idx = offs[lo:hi, None] + np.arange(64)[None, :]
        blob[idx.ravel()] = hb.ravel()
        # payloads
        for k, c in enumerate(classes):
            m = np.flatnonzero(c[lo:hi]) + lo
            if m.size == 0:
                continue
            blobk = pools[k][0]
            ln = pay_len[m]
            tot = int(ln.sum())
            rep = np.repeat(np.arange(m.size), ln)
            within = np.arange(tot) - np.repeat(np.cumsum(ln) - ln, ln)
            dsti = (offs[m] + hdr_len[m])[rep] + within
            srci = pay_src[m][rep] + within
            # payload overwrites the zero padding / tail of the 64 B header template
            blob[dsti] = blobk[srci]
    info = {"classes": np.bincount(cls, minlength=len(mix)).tolist(), "bytes": int(offs[-1]), "pool": pool}
    return blob, (offs // 16).astype(np.uint32), info


# ---------------------------------------------------------------------------------------------
# pcap (classic libpcap, LINKTYPE_ETHERNET) — the on-wire file format either side of the path
# ---------------------------------------------------------------------------------------------

def write_pcap(path: str, frames: list[bytes]) -> None:
    with open(path, "wb") as fh:
        fh.write(struct.pack("<IHHiIII", 0xA1B2C3D4, 2, 4, 0, 0, 65535, 1))
        for i, fr in enumerate(frames):
            fh.write(struct.pack("<IIII", i // 1000000, i % 1000000, len(fr), len(fr)) + fr)


def read_pcap(path: str) -> tuple[np.ndarray, np.ndarray]:
    """Read a classic pcap into the varlen layout (blob, offs16)."""
    data = open(path, "rb").read()
    magic = struct.unpack_from("<I", data, 0)[0]
    if magic not in (0xA1B2C3D4, 0xA1B23C4D):
        raise ValueError("unsupported pcap magic %#x" % magic)
    frames, o = [], 24
    while o + 16 <= len(data):
        caplen = struct.unpack_from("<I", data, o + 8)[0]
        frames.append(data[o + 16:o + 16 + caplen])
        o += 16 + caplen
    return build_varlen(frames)
