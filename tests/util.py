"""Shared helpers for the parity tests (tests are the only place product and oracle meet)."""
from __future__ import annotations

import os

import numpy as np

from clawker_b200 import abi, capture, engine
from oracle import oracle, rules_oracle


def oracle_tables(w: capture.Workload, with_cidrs: bool = True) -> dict:
    """Neutral table set for the oracle: base tables + what the (Python-restated) rule compiler derives."""
    c = rules_oracle.compile_rules(w.rules)
    t = dict(w.tables)
    t["zones"], t["snis"], t["routes"], t["http"] = c.zones, c.snis, dict(c.routes), list(c.http)
    t["cidrs"] = list(c.cidrs) if with_cidrs else []
    return t


def make_engine(w: capture.Workload, loader: int | None = None, with_cidrs: bool = True, **kw) -> engine.Engine:
    if loader is not None:
        os.environ["CLK_LOADER"] = str(loader)
    else:
        os.environ.pop("CLK_LOADER", None)
    kw.setdefault("max_metrics", 1 << 20)
    kw.setdefault("max_dns_cache", 1 << 21)
    kw.setdefault("max_routes", 1 << 16)
    kw.setdefault("max_containers", 4096)
    e = engine.Engine(**kw)
    e.load_tables(w.tables)
    rules = w.rules if with_cidrs else [r for r in w.rules if not (r["dst"][0].isdigit() and "/" in r["dst"])]
    e.compile_rules(rules)
    e.publish()
    return e


def assert_same_verdicts(got: np.ndarray, want: np.ndarray, what: str = ""):
    if np.array_equal(got, want):
        return
    bad = np.flatnonzero(got != want)
    i = int(bad[0])
    f = abi.verdict_fields
    raise AssertionError(f"{what}: {bad.size}/{got.size} verdicts differ; first at {i}: "
                         f"got {got[i]:#010x} {f(int(got[i]))} want {want[i]:#010x} {f(int(want[i]))}")


def assert_same_metrics(got: np.ndarray, want: np.ndarray, what: str = ""):
    g, w = oracle.sort_metrics(got.copy()), oracle.sort_metrics(want.copy())
    g["_pad"] = 0
    w["_pad"] = 0
    if g.size != w.size or not np.array_equal(g, w):
        raise AssertionError(f"{what}: metrics differ: {g.size} rows vs {w.size}; total {int(g['count'].sum())} vs {int(w['count'].sum())}")
