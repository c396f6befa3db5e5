"""The `ebpf-manager` verbs (internal/ebpf/cmd/main.go:30-303) on the GPU engine's pinned state: same argv, same
JSON payloads, same messages and exit codes; the state file they leave equals what driving ebpf.Manager directly
gives.  No GPU needed: the control-plane binary runs on a host-only engine."""
import json
import os
import subprocess

import pytest

from clawker_b200 import abi, ebpf, engine

CFG = {"envoy_ip": "172.18.0.2", "coredns_ip": "172.18.0.3", "gateway_ip": "172.18.0.1", "cidr": "172.18.0.0/16",
       "host_proxy_ip": "192.168.65.254", "host_proxy_port": 18374, "egress_port": 10000}


@pytest.fixture
def box(tmp_path):
    engine.build_library()
    root = tmp_path / "cgroup"
    for name in ("docker-aaa.scope", "docker-bbb.scope"):
        (root / "system.slice" / name).mkdir(parents=True)
    env = dict(os.environ, CLK_PIN_PATH=str(tmp_path / "bpf" / "maps.pin"), CLK_CGROUP_ROOT=str(root))

    def run(*argv, ok=True):
        r = subprocess.run([str(engine.CLI_PATH), *argv], env=env, capture_output=True, text=True, timeout=60)
        assert (r.returncode == 0) == ok, (argv, r.returncode, r.stdout, r.stderr)
        return r

    run.pin = env["CLK_PIN_PATH"]
    run.cg = lambda n: str(root / "system.slice" / f"docker-{n}.scope")
    run.ino = lambda n: os.stat(run.cg(n)).st_ino
    return run


def _state(path):
    e = engine.Engine(host_only=True)
    e.state_load(path)
    return e


def test_verbs_match_manager(box):
    assert box("init").stdout == "eBPF programs loaded and pinned\n"
    assert box("enable", box.cg("aaa"), json.dumps(CFG)).stdout == f"enabled cgroup_id={box.ino('aaa')}\n"
    assert box("enable", box.cg("bbb"), json.dumps(dict(CFG, host_proxy_ip="", container_ip="172.18.0.9", uid=501))).stdout.startswith("enabled")
    assert box("bypass", box.cg("bbb")).stdout == f"bypass enabled cgroup_id={box.ino('bbb')}\n"
    routes = [{"domain_hash": ebpf.DomainHash("github.com"), "dst_port": 22, "envoy_port": 10001},
              {"domain_hash": 7, "dst_port": 8080, "envoy_port": 10002}]
    assert box("sync-routes", json.dumps(routes)).stdout == "synced 2 routes\n"
    assert box("dns-update", "140.82.112.3", str(ebpf.DomainHash("github.com")), "300").stdout == ""
    assert box("dns-update", "10.9.9.9", "5", "0").stdout == ""
    assert box("gc-dns").stdout == "gc-dns: removed 0 expired entries\n"        # ttl 0 expires at `now`: not < now yet

    want = ebpf.Manager(engine.Engine(host_only=True))
    cfg = ebpf.NewContainerConfig(CFG["envoy_ip"], CFG["coredns_ip"], CFG["gateway_ip"], CFG["cidr"], CFG["host_proxy_ip"], 18374, 10000)
    want.Enable(box.ino("aaa"), box.cg("aaa"), cfg)
    want.Enable(box.ino("bbb"), box.cg("bbb"), ebpf.NewContainerConfig(CFG["envoy_ip"], CFG["coredns_ip"], CFG["gateway_ip"], CFG["cidr"], "", 18374, 10000))
    want.BindEndpoint("172.18.0.9", box.ino("bbb"), uid=501)
    want.Bypass(box.ino("bbb"))
    want.SyncRoutes([ebpf.Route(r["domain_hash"], r["dst_port"], r["envoy_port"]) for r in routes])
    got = _state(box.pin)
    for m in (abi.MAP_CONTAINER, abi.MAP_BYPASS, abi.MAP_ROUTE, abi.MAP_ENDPOINT):
        assert got.map_rows(m) == want.e.map_rows(m), m
    dns = got.map_rows(abi.MAP_DNS_CACHE)
    assert len(dns) == 2 and dns[bytes([140, 82, 112, 3])][:4] == ebpf.DomainHash("github.com").to_bytes(4, "little")

    out = box("dump", box.cg("aaa")).stdout
    assert out == (f"cgroup_id={box.ino('aaa')}\n  envoy_ip=172.18.0.2 coredns_ip=172.18.0.3 gateway_ip=172.18.0.1\n"
                   "  net_addr=172.18.0.0 net_mask=255.255.0.0\n  host_proxy_ip=192.168.65.254 host_proxy_port=18374 egress_port=10000\n")
    assert box("unbypass", box.cg("bbb")).stdout == f"bypass disabled cgroup_id={box.ino('bbb')}\n"
    assert box("enable", box.cg("bbb"), json.dumps(CFG)).returncode == 0           # Enable clears a stale bypass flag
    assert box("disable", box.cg("aaa")).stdout == f"disabled cgroup_id={box.ino('aaa')}\n"
    assert box("disable", box.cg("aaa")).returncode == 0                           # ErrKeyNotExist is success
    got = _state(box.pin)
    assert got.map_count(abi.MAP_CONTAINER) == 1 and got.map_count(abi.MAP_BYPASS) == 0
    assert "no container_map entry" in box("dump", box.cg("aaa")).stdout
    assert box("sync-routes", "[]").stdout == "synced 0 routes\n" and _state(box.pin).map_count(abi.MAP_ROUTE) == 0


def test_errors_exit_1_with_the_reference_prefix(box):
    assert box(ok=False).returncode == 1                                            # no verb: usage
    r = box("frobnicate", ok=False)
    assert "unknown command: frobnicate" in r.stderr and "Usage: ebpf-manager" in r.stderr
    assert "run `ebpf-manager init` first" in box("gc-dns", ok=False).stderr        # OpenPinned before Load
    box("init")
    assert "Usage" in box("enable", box.cg("aaa"), ok=False).stderr                 # requireArgs
    assert box("enable", box.cg("aaa"), "{not json", ok=False).stderr.startswith("ebpf-manager enable: parsing config JSON:")
    assert "parsing CIDR nope" in box("enable", box.cg("aaa"), json.dumps(dict(CFG, cidr="nope")), ok=False).stderr
    assert "parsing config JSON" in box("enable", box.cg("aaa"), json.dumps(dict(CFG, egress_port=70000)), ok=False).stderr
    assert "parsing config JSON" in box("enable", box.cg("aaa"), json.dumps(dict(CFG, egress_port="10000")), ok=False).stderr
    for bad, msg in (("", "cgroup path is empty"), ("/etc/passwd", "must be under"), (box.cg("aaa") + "/../docker-bbb.scope", "must not contain '..'"),
                     (box.cg("zzz"), "opening cgroup")):
        assert msg in box("disable", bad, ok=False).stderr, bad
    assert "invalid IP: 1.2.3" in box("dns-update", "1.2.3", "5", "60", ok=False).stderr
    assert "parsing domain hash" in box("dns-update", "1.2.3.4", "-5", "60", ok=False).stderr
    assert "parsing TTL" in box("dns-update", "1.2.3.4", "5", "4294967296", ok=False).stderr
    assert "parsing routes JSON" in box("sync-routes", '[{"domain_hash": "x"}]', ok=False).stderr
    assert _state(box.pin).map_count(abi.MAP_CONTAINER) == 0                        # failed verbs leave the state alone


def test_state_file_roundtrip_and_rejects_garbage(tmp_path):
    e = ebpf.Manager(engine.Engine(host_only=True))
    e.Enable(42, "/x", ebpf.NewContainerConfig("10.0.0.2", "10.0.0.3", "10.0.0.1", "10.0.0.0/24", "", 0, 10000))
    e.e.compile_rules([{"dst": "github.com", "proto": "tls", "port": 443, "action": "allow"},
                       {"dst": "github.com", "proto": "tcp", "port": 22, "action": "allow"}, {"dst": "10.1.0.0/16", "action": "allow"}])
    e.e.map_update(abi.MAP_ENDPOINT6, (abi.C.c_uint8 * 16)(*range(16)), abi.Endpoint(42, 1000, 0))
    p = str(tmp_path / "maps.pin")
    e.e.state_save(p)
    g = _state(p)
    for m in (abi.MAP_CONTAINER, abi.MAP_ROUTE, abi.MAP_ENDPOINT6, abi.MAP_CIDR):
        assert g.map_rows(m) == e.e.map_rows(m) and len(g.map_rows(m)) >= 1
    for kind in (0, 1, 2):
        assert g.rules_dump(kind) == e.e.rules_dump(kind)
    raw = open(p, "rb").read()
    for bad in (b"", b"nonsense" * 8, raw[:-9], raw[:40]):
        q = str(tmp_path / "bad.pin")
        open(q, "wb").write(bad)
        with pytest.raises(engine.ClkError):
            g.state_load(q)
        assert g.map_rows(abi.MAP_CONTAINER) == e.e.map_rows(abi.MAP_CONTAINER)      # a bad file leaves the tables untouched
    with pytest.raises(engine.ClkError):
        g.state_load(str(tmp_path / "missing.pin"))
    small = engine.Engine(host_only=True, max_routes=0, max_containers=0)
    small.state_load(p)                                                             # 0 = reference defaults: fits


def test_state_loader_survives_corruption(tmp_path):
    """Random truncations and byte flips of a valid state file: an error or a clean load, never a crash or a hang."""
    import numpy as np
    e = ebpf.Manager(engine.Engine(host_only=True))
    for cg in range(1, 20):
        e.Enable(cg, "/x", ebpf.NewContainerConfig("10.0.0.2", "10.0.0.3", "10.0.0.1", "10.0.0.0/24", "", 0, 10000))
        e.BindEndpoint("10.0.0.%d" % (cg + 10), cg)
    e.e.compile_rules([{"dst": "d%d.example.com" % i, "proto": ["tls", "tcp", "http"][i % 3], "port": [443, 22, 80][i % 3],
                        "path_rules": [("/a", "allow")] if i % 3 == 2 else []} for i in range(40)] + [{"dst": "10.%d.0.0/16" % i} for i in range(10)])
    p = str(tmp_path / "maps.pin")
    e.e.state_save(p)
    raw = bytearray(open(p, "rb").read())
    rng = np.random.Generator(np.random.PCG64(5))
    g = engine.Engine(host_only=True)
    ok = bad = 0
    for _ in range(300):
        b = bytearray(raw)
        if rng.random() < 0.5:
            b = b[: int(rng.integers(0, len(b)))]
        for _k in range(int(rng.integers(0, 6))):
            if b:
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        q = str(tmp_path / "fuzz.pin")
        open(q, "wb").write(bytes(b))
        try:
            g.state_load(q)
            ok += 1
        except engine.ClkError:
            bad += 1
    assert ok + bad == 300 and bad > 100
    g.state_load(p)                                                                  # and the intact file still loads
    assert g.map_rows(abi.MAP_CONTAINER) == e.e.map_rows(abi.MAP_CONTAINER)


# TestValidateCgroupPath / TestCgroupID_RejectsMaliciousPath (internal/ebpf/manager_test.go:99-189,363-420), as data.
# (The NUL-byte case cannot travel through argv; exec refuses it before the binary runs.)
CGROUP_PATH_VECTORS = [
    ("/sys/fs/cgroup/system.slice/docker-abc123.scope", None, "/sys/fs/cgroup/system.slice/docker-abc123.scope"),
    ("/sys/fs/cgroup/docker/abc123", None, "/sys/fs/cgroup/docker/abc123"),
    ("/sys/fs/cgroup//system.slice/./docker-abc.scope", None, "/sys/fs/cgroup/system.slice/docker-abc.scope"),
    ("", "empty", None),
    ("/sys/fs/cgroup/../etc/passwd", "'..'", None),
    ("/sys/fs/cgroup/../etc/shadow", "'..'", None),
    ("../../etc/passwd", "'..'", None),
    ("/etc/passwd", "under /sys/fs/cgroup/", None),
    ("/proc/self/environ", "under /sys/fs/cgroup/", None),
    ("cgroup/foo", "under /sys/fs/cgroup/", None),
    ("/sys/fs/cgroup", "under /sys/fs/cgroup/", None),
    ("/sys/fs/cgroup/system.slice/x\ny", "illegal characters", None),
]


@pytest.mark.parametrize("path,want_err,cleaned", CGROUP_PATH_VECTORS)
def test_cgroup_path_validation_reference_vectors(tmp_path, path, want_err, cleaned):
    engine.build_library()
    env = {k: v for k, v in os.environ.items() if k != "CLK_CGROUP_ROOT"}
    env["CLK_PIN_PATH"] = str(tmp_path / "maps.pin")
    r = subprocess.run([str(engine.CLI_PATH), "disable", path], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1
    if want_err:
        assert want_err in r.stderr, r.stderr
    else:                                   # passes validation; the (cleaned) path simply does not exist on this box
        assert f"opening cgroup: open {cleaned}:" in r.stderr, r.stderr
