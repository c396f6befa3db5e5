"""metrics_map exporter: Prometheus text, bounded cardinality, deltas between scrapes."""
import urllib.request

import numpy as np

from clawker_b200 import ebpf, exporter
from clawker_b200.abi import METRIC_DTYPE, TOTAL_DTYPE


def _rows(*t):
    a = np.zeros(len(t), dtype=METRIC_DTYPE)
    for i, (cg, h, port, act, c) in enumerate(t):
        a[i] = (cg, h, port, act, 0, c)
    return a


class FakeEngine:
    def __init__(self, rows):
        self.rows = rows

    def counters(self, merged=False):
        return self.rows


GH = ebpf.DomainHash("github.com")


def test_render_labels_and_order():
    rows = _rows((7, GH, 443, 0, 12), (7, 0, 8080, 1, 3), (9, 555, 22, 2, 1), (7, GH, 443, 0, 5))
    tot = np.zeros(1, dtype=TOTAL_DTYPE)
    tot[0] = (7, (17, 3, 0))
    txt = exporter.render(rows, domains={GH: "github.com"}, containers={7: 'agent "a"'}, totals=tot)
    lines = [l for l in txt.splitlines() if not l.startswith("#")]
    assert lines[0] == 'clawker_egress_decisions_total{cgroup_id="7",container="agent \\"a\\"",domain="-",port="8080",action="deny"} 3'
    assert f'domain="github.com",port="443",action="allow"}} 17' in lines[1]            # duplicate keys (two GPUs) are summed
    assert lines[2] == 'clawker_egress_decisions_total{cgroup_id="9",domain="hash:555",port="22",action="bypass"} 1'
    assert 'clawker_egress_container_decisions_total{cgroup_id="7",container="agent \\"a\\"",action="allow"} 17' in lines
    assert txt.count("# TYPE") == 2 and txt.endswith("\n")


def test_cardinality_collapse_keeps_the_largest_domains():
    rows = _rows((1, 10, 80, 0, 100), (1, 11, 80, 0, 50), (1, 12, 80, 1, 5), (1, 13, 443, 1, 4), (2, 10, 80, 0, 1))
    txt = exporter.render(rows, max_domains_per_container=2)
    assert 'cgroup_id="1",domain="_other",port="0",action="deny"} 9' in txt
    assert 'domain="hash:10"' in txt and 'domain="hash:11"' in txt and 'hash:12' not in txt and 'hash:13' not in txt
    total = sum(int(l.rsplit(" ", 1)[1]) for l in txt.splitlines() if not l.startswith("#"))
    assert total == int(rows["count"].sum())                                           # nothing lost by collapsing


def test_scraper_deltas_and_reset():
    eng = FakeEngine(_rows((1, 10, 80, 0, 100), (1, 11, 80, 1, 7)))
    s = exporter.Scraper(eng)
    assert s.deltas() == {(1, 10, 80, 0): 100, (1, 11, 80, 1): 7}
    eng.rows = _rows((1, 10, 80, 0, 130), (1, 11, 80, 1, 7), (2, 0, 53, 0, 2))
    assert s.deltas() == {(1, 10, 80, 0): 30, (2, 0, 53, 0): 2}
    eng.rows = _rows((1, 10, 80, 0, 4))                                                 # engine reset: counter restarts
    assert s.deltas() == {(1, 10, 80, 0): 4}


def test_metrics_endpoint():
    srv = exporter.serve(FakeEngine(_rows((3, 0, 53, 0, 9))), port=0)
    try:
        port = srv.server_address[1]
        body = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=10).read().decode()
        assert 'cgroup_id="3",domain="-",port="53",action="allow"} 9' in body
        try:
            urllib.request.urlopen(f"http://127.0.0.1:{port}/other", timeout=10)
            assert False
        except urllib.error.HTTPError as ex:
            assert ex.code == 404
    finally:
        srv.shutdown()
