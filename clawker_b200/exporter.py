"""metrics_map scraper / exporter.

The reference's programs count every decision in `metrics_map` (`metric_inc`, internal/ebpf/bpf/common.h:190-210)
but nothing reads it: the scraper is a planned feature (.serena/memories/outstanding-features.md:262-300 — a
`ScrapeMetrics()` reader on an interval, deltas, labels container / domain / port / action, a Prometheus `/metrics`
endpoint like Envoy's and CoreDNS's, cardinality collapsed above N domains per container).  This module is that
reader over the GPU engine's counters: `Engine.counters()` for one GPU, `Engine.counters(merged=True)` after the
per-key merge, `Engine.totals_merge()` for the per-batch per-container totals."""
from __future__ import annotations

import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import numpy as np

ACTIONS = ("allow", "deny", "bypass")                     # ebpf.Action types.go:76-80
METRIC = "clawker_egress_decisions_total"
TOTAL = "clawker_egress_container_decisions_total"


def _esc(v) -> str:
    return str(v).replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n")


def _key(r):
    return int(r["cgroup_id"]), int(r["domain_hash"]), int(r["dst_port"]), int(r["action"])


class Scraper:
    """ScrapeMetrics(): the current rows, and what changed since the previous scrape (counters only go up; a row that
    went away — engine reset — restarts from zero, as a Prometheus counter reset does)."""

    def __init__(self, engine, merged: bool = False):
        self.e, self.merged, self.last = engine, merged, {}

    def scrape(self) -> np.ndarray:
        return self.e.counters(merged=self.merged)

    def deltas(self, rows: np.ndarray | None = None) -> dict:
        rows = self.scrape() if rows is None else rows
        cur = {_key(r): int(r["count"]) for r in rows}
        out = {k: v - self.last.get(k, 0) if v >= self.last.get(k, 0) else v for k, v in cur.items()}
        self.last = cur
        return {k: v for k, v in out.items() if v}


def render(rows: np.ndarray, domains: dict | None = None, containers: dict | None = None,
           max_domains_per_container: int = 0, totals: np.ndarray | None = None) -> str:
    """Prometheus text exposition.  `domains`: DomainHash -> name (the rule set knows them; unknown hashes print as
    hash:<n>, 0 as "-"); `containers`: cgroup id -> container name.  With max_domains_per_container > 0 the smallest
    domains of a container beyond that many are summed into domain="_other" (bounded cardinality)."""
    domains, containers = domains or {}, containers or {}
    agg: dict = {}
    for r in rows:
        cg, h, port, act = _key(r)
        agg[(cg, h, port, act)] = agg.get((cg, h, port, act), 0) + int(r["count"])
    if max_domains_per_container > 0:
        per: dict = {}
        for (cg, h, _, _), c in agg.items():
            per.setdefault(cg, {}).setdefault(h, 0)
            per[cg][h] += c
        keep = {cg: set(sorted(d, key=lambda h: (-d[h], h))[:max_domains_per_container]) for cg, d in per.items()}
        folded: dict = {}
        for (cg, h, port, act), c in agg.items():
            k = (cg, h, port, act) if h in keep[cg] else (cg, None, 0, act)
            folded[k] = folded.get(k, 0) + c
        agg = folded
    out = [f"# HELP {METRIC} Egress decisions counted by the classifier (metrics_map).", f"# TYPE {METRIC} counter"]
    for (cg, h, port, act) in sorted(agg, key=lambda k: (k[0], -1 if k[1] is None else k[1], k[2], k[3])):
        dom = "_other" if h is None else "-" if h == 0 else domains.get(h, f"hash:{h}")
        lab = [f'cgroup_id="{cg}"']
        if cg in containers:
            lab.append(f'container="{_esc(containers[cg])}"')
        lab += [f'domain="{_esc(dom)}"', f'port="{port}"', f'action="{ACTIONS[act] if act < 3 else act}"']
        out.append(f"{METRIC}{{{','.join(lab)}}} {agg[(cg, h, port, act)]}")
    if totals is not None:
        out += [f"# HELP {TOTAL} Per-container sums over all GPUs (per-batch all-gather).", f"# TYPE {TOTAL} counter"]
        for t in totals:
            for a in range(3):
                lab = [f'cgroup_id="{int(t["cgroup_id"])}"']
                if int(t["cgroup_id"]) in containers:
                    lab.append(f'container="{_esc(containers[int(t["cgroup_id"])])}"')
                out.append(f"{TOTAL}{{{','.join(lab)},action=\"{ACTIONS[a]}\"}} {int(t['count'][a])}")
    return "\n".join(out) + "\n"


def serve(engine, host: str = "127.0.0.1", port: int = 9419, merged: bool = False, **render_kw) -> ThreadingHTTPServer:
    """/metrics endpoint over the engine's counters (started on a daemon thread; .shutdown() stops it)."""
    lock = threading.Lock()

    class H(BaseHTTPRequestHandler):
        def do_GET(self):
            if self.path.split("?")[0] != "/metrics":
                self.send_error(404)
                return
            with lock:
                body = render(engine.counters(merged=merged), **render_kw).encode()
            self.send_response(200)
            self.send_header("Content-Type", "text/plain; version=0.0.4")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *a):
            pass

    srv = ThreadingHTTPServer((host, port), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    return srv
