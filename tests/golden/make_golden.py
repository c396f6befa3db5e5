"""Generate tests/golden/events_ref.npz: socket events + the verdict words and metric rows produced by the
REFERENCE's own eBPF programs (oracle/_ref/liboracle_ref.so = /root/reference/internal/ebpf/bpf/clawker.c
compiled for the host).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixture travels to the GPU box (which has no /root/reference); tests compare the C restatement and the
CUDA path against it.  Also writes packets_ref.npz: 4096 64-byte slots + reference verdicts (no CIDR rules).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from clawker_b200 import abi, capture  # noqa: E402
from oracle import oracle  # noqa: E402
from util import oracle_tables  # noqa: E402


def exhaustive_events(w):
    """Every hook x socket type x identity class x destination class (grid, not random)."""
    ip4 = capture.ip4
    eps = w.tables["endpoints"]
    idents = list(eps.values()) + [(999999, 1000, 0), (1001, 0, 0), (1001, 1000, 0xC1A4)]
    dsts = [ip4("127.0.0.1"), ip4("127.0.0.11"), ip4("172.20.0.5"), ip4(capture.GATEWAY_IP), ip4(capture.HOST_PROXY_IP),
            ip4(capture.COREDNS_IP), ip4(capture.ENVOY_IP), ip4("8.8.8.8"), ip4("0.0.0.0"), 0xFFFFFFFF] + list(w.tables["dns_cache"])[:6]
    route_ports = sorted({r["port"] for r in w.rules if r["proto"] in ("tcp", "ssh")})[:3]
    ports = [53, 443, 80, 22, capture.HOST_PROXY_PORT, 10000, 0, 65535] + route_ports
    rows = []
    for hook in range(8):
        for st in (1, 2, 3, 5):
            for cg, uid, mark in idents:
                for d in dsts:
                    for p in ports:
                        for k6 in ((0, 0, 0xFFFF0000, d), (0, 0, 0, 0x01000000), (0x20, 1, 2, d)) if 4 <= hook <= 6 else ((0, 0, 0, 0),):
                            rows.append((cg, uid, mark, d, k6, p, hook, st))
    ev = np.zeros(len(rows), dtype=abi.EVENT_DTYPE)
    for i, (cg, uid, mark, d, k6, p, hook, st) in enumerate(rows):
        ev[i] = (cg, uid, mark, d, k6, p, hook, st)
    return ev


def main():
    assert oracle.have_ref(), "oracle/_ref/liboracle_ref.so missing: run `make -C oracle` where /root/reference exists"
    w = capture.make_workload("c2")
    t = oracle_tables(w, with_cidrs=False)
    ev = np.concatenate([exhaustive_events(w), capture.gen_events(w, 20000, seed=9)])
    verd, met = oracle.Oracle("ref").load(t).classify_events(ev)
    out = Path(__file__).resolve().parent
    np.savez_compressed(out / "events_ref.npz", events=ev, verdicts=verd, metrics=oracle.sort_metrics(met))
    pk = capture.gen_slots(w, 4096, seed=77)
    pv, pm = oracle.Oracle("ref").load(t).classify(pk)
    np.savez_compressed(out / "packets_ref.npz", packets=pk, verdicts=pv, metrics=oracle.sort_metrics(pm))
    print(f"events_ref.npz: {ev.size} events, {met.size} metric rows; packets_ref.npz: {pk.shape[0]} packets")


if __name__ == "__main__":
    main()
