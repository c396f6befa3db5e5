"""ctypes / numpy mirrors of include/clawker_gpu.h (row layouts of internal/ebpf/bpf/common.h:80-169)."""
from __future__ import annotations

import ctypes as C

import numpy as np


class ContainerConfig(C.Structure):            # clk_container_config / struct container_config common.h:80-89
    _fields_ = [("envoy_ip", C.c_uint32), ("coredns_ip", C.c_uint32), ("gateway_ip", C.c_uint32),
                ("net_addr", C.c_uint32), ("net_mask", C.c_uint32), ("host_proxy_ip", C.c_uint32),
                ("host_proxy_port", C.c_uint16), ("egress_port", C.c_uint16)]


class DNSEntry(C.Structure):                   # struct dns_entry common.h:96-99
    _fields_ = [("domain_hash", C.c_uint32), ("expire_ts", C.c_uint32)]


class RouteKey(C.Structure):                   # struct route_key common.h:103-107
    _fields_ = [("domain_hash", C.c_uint32), ("dst_port", C.c_uint16), ("_pad", C.c_uint16)]


class RouteVal(C.Structure):                   # struct route_val common.h:110-113
    _fields_ = [("envoy_port", C.c_uint16), ("_pad", C.c_uint16)]


class Route(C.Structure):                      # ebpf.Route manager.go:466-470
    _fields_ = [("domain_hash", C.c_uint32), ("dst_port", C.c_uint16), ("envoy_port", C.c_uint16)]


class MetricKey(C.Structure):                  # struct metric_key common.h:163-169
    _fields_ = [("cgroup_id", C.c_uint64), ("domain_hash", C.c_uint32), ("dst_port", C.c_uint16),
                ("action", C.c_uint8), ("_pad", C.c_uint8)]


class Endpoint(C.Structure):
    _fields_ = [("cgroup_id", C.c_uint64), ("uid", C.c_uint32), ("so_mark", C.c_uint32)]


class Cidr(C.Structure):
    _fields_ = [("addr", C.c_uint32), ("mask", C.c_uint32)]


class Cfg(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("max_containers", C.c_uint32),
                ("max_dns_cache", C.c_uint32), ("max_routes", C.c_uint32), ("max_metrics", C.c_uint32),
                ("max_cidrs", C.c_uint32), ("flags", C.c_uint32), ("reserved_sms", C.c_uint32)]


class PathRule(C.Structure):                   # config.PathRule schema.go:124-127
    _fields_ = [("path", C.c_char_p), ("action", C.c_char_p)]


class Rule(C.Structure):                       # config.EgressRule schema.go:131-138
    _fields_ = [("dst", C.c_char_p), ("proto", C.c_char_p), ("port", C.c_int32), ("action", C.c_char_p),
                ("path_default", C.c_char_p), ("path_rules", C.POINTER(PathRule)), ("n_path_rules", C.c_uint32)]


class Ports(C.Structure):
    _fields_ = [("egress_port", C.c_int32), ("tcp_port_base", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("last_kernel_ms", C.c_float), ("sm_count", C.c_uint32),
                ("last_grid", C.c_uint32), ("last_block", C.c_uint32), ("last_smem_bytes", C.c_uint32),
                ("metrics_dropped", C.c_uint64)]


EVENT_DTYPE = np.dtype([("cgroup_id", "<u8"), ("uid", "<u4"), ("so_mark", "<u4"), ("user_ip4", "<u4"),
                        ("user_ip6", "<u4", (4,)), ("user_port", "<u2"), ("hook", "u1"), ("sock_type", "u1")])
METRIC_DTYPE = np.dtype([("cgroup_id", "<u8"), ("domain_hash", "<u4"), ("dst_port", "<u2"), ("action", "u1"),
                         ("_pad", "u1"), ("count", "<u8")])
TOTAL_DTYPE = np.dtype([("cgroup_id", "<u8"), ("count", "<u8", (3,))])           # clk_total: ALLOW, DENY, BYPASS
HDR16_DTYPE = np.dtype([("src", "<u4"), ("dst", "<u4"), ("sport", "<u2"), ("dport", "<u2"), ("proto", "u1"), ("kind", "u1"), ("_pad", "<u2")])   # clk_hdr16
assert HDR16_DTYPE.itemsize == 16
DNSUPD_DTYPE = np.dtype([("ip", "<u4"), ("domain_hash", "<u4"), ("ttl", "<u4"), ("packet", "<u4")])
assert EVENT_DTYPE.itemsize == 40 and METRIC_DTYPE.itemsize == 24 and DNSUPD_DTYPE.itemsize == 16 and TOTAL_DTYPE.itemsize == 32

MAP_CONTAINER, MAP_BYPASS, MAP_DNS_CACHE, MAP_ROUTE, MAP_METRICS, MAP_ENDPOINT, MAP_ENDPOINT6, MAP_CIDR = range(8)
F_NAME_STAGE, F_NO_METRICS = 1, 2
OK, ENOENT, EIO, E2BIG, ENOMEM, EEXIST, EINVAL, ENOSYS = 0, -2, -5, -7, -12, -17, -22, -38


def verdict_fields(w):
    """Split verdict words (numpy array or int) into named fields (include/clawker_gpu.h)."""
    return dict(verdict=w & 3, metric=(w >> 2) & 3, name=(w >> 4) & 3, target=(w >> 6) & 3,
                port=(w >> 8) & 0xFFFF, hook=(w >> 24) & 7, nupd=(w >> 27) & 31)
