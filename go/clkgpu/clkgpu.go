// Package clkgpu is the cgo shim over libclawker_gpu.so (include/clawker_gpu.h).
//
// It contains no logic: every function marshals Go values into one clk_* call and maps the
// negative return code to a Go error.  No Go pointer is retained by C across calls (cgo rule);
// packet and verdict slices are pinned only for the duration of the call by cgo itself.
package clkgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -lclawker_gpu
#include <stdlib.h>
#include "clawker_gpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"strings"
	"unsafe"
)

// ErrKeyNotExist mirrors ebpf.ErrKeyNotExist (CLK_ENOENT).
var ErrKeyNotExist = errors.New("key does not exist")

// ErrMapFull mirrors the kernel's E2BIG on a full BPF hash map (CLK_E2BIG).
var ErrMapFull = errors.New("map full")

// Map ids (CLK_MAP_*).
const (
	MapContainer = C.CLK_MAP_CONTAINER
	MapBypass    = C.CLK_MAP_BYPASS
	MapDNSCache  = C.CLK_MAP_DNS_CACHE
	MapRoute     = C.CLK_MAP_ROUTE
	MapEndpoint  = C.CLK_MAP_ENDPOINT
	MapEndpoint6 = C.CLK_MAP_ENDPOINT6
	MapCIDR      = C.CLK_MAP_CIDR
)

// Classify flags (CLK_F_*).
const (
	FlagNameStage = C.CLK_F_NAME_STAGE
	FlagNoMetrics = C.CLK_F_NO_METRICS
)

// Engine is one GPU's classifier instance.
type Engine struct{ h *C.clk_engine }

// Config mirrors clk_cfg; zero values select the reference's max_entries.
type Config struct {
	Device                                                         int
	MaxContainers, MaxDNSCache, MaxRoutes, MaxMetrics, MaxCIDRs uint32
	ReservedSMs                                                    uint32 // SMs left to the side-stream collective (multi-GPU hosts)
}

func (e *Engine) err(rc C.int) error {
	switch rc {
	case C.CLK_OK:
		return nil
	case C.CLK_ENOENT:
		return ErrKeyNotExist
	case C.CLK_E2BIG:
		return fmt.Errorf("%w: %s", ErrMapFull, C.GoString(C.clk_last_error(e.h)))
	default:
		return fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_last_error(e.h)), int(rc))
	}
}

// New creates the tables on one GPU (replaces loading and pinning the BPF objects).
func New(cfg Config) (*Engine, error) {
	c := C.clk_cfg{abi_version: C.CLK_ABI_VERSION, device: C.int32_t(cfg.Device),
		max_containers: C.uint32_t(cfg.MaxContainers), max_dns_cache: C.uint32_t(cfg.MaxDNSCache),
		max_routes: C.uint32_t(cfg.MaxRoutes), max_metrics: C.uint32_t(cfg.MaxMetrics), max_cidrs: C.uint32_t(cfg.MaxCIDRs), reserved_sms: C.uint32_t(cfg.ReservedSMs)}
	var h *C.clk_engine
	if rc := C.clk_engine_create(&c, &h); rc != C.CLK_OK {
		return nil, fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_last_error(nil)), int(rc))
	}
	return &Engine{h: h}, nil
}

// Close releases the engine.
func (e *Engine) Close() error {
	if e.h != nil {
		C.clk_engine_destroy(e.h)
		e.h = nil
	}
	return nil
}

// MapUpdate / MapDelete / MapLookup take the reference's row structs by pointer and size.
func (e *Engine) MapUpdate(id int, key unsafe.Pointer, klen uintptr, val unsafe.Pointer, vlen uintptr) error {
	return e.err(C.clk_map_update(e.h, C.int(id), key, C.size_t(klen), val, C.size_t(vlen)))
}

func (e *Engine) MapDelete(id int, key unsafe.Pointer, klen uintptr) error {
	return e.err(C.clk_map_delete(e.h, C.int(id), key, C.size_t(klen)))
}

func (e *Engine) MapLookup(id int, key unsafe.Pointer, klen uintptr, val unsafe.Pointer, vlen uintptr) error {
	return e.err(C.clk_map_lookup(e.h, C.int(id), key, C.size_t(klen), val, C.size_t(vlen)))
}

// Route is clk_route / ebpf.Route.
type Route struct {
	DomainHash uint32
	DstPort    uint16
	EnvoyPort  uint16
}

// RoutesReplace is Manager.SyncRoutes as one atomic replace.
func (e *Engine) RoutesReplace(routes []Route) error {
	var p *C.clk_route
	if len(routes) > 0 {
		p = (*C.clk_route)(unsafe.Pointer(&routes[0]))
	}
	return e.err(C.clk_routes_replace(e.h, p, C.uint32_t(len(routes))))
}

// DNSUpdate is dns_cache[ip] = {hash, now+ttl}; now == 0 means time.Now().
func (e *Engine) DNSUpdate(ip, domainHash, ttl, now uint32) error {
	return e.err(C.clk_dns_update(e.h, C.uint32_t(ip), C.uint32_t(domainHash), C.uint32_t(ttl), C.uint32_t(now)))
}

// DNSGC removes expired dns_cache rows and returns how many were cleared.
func (e *Engine) DNSGC(now uint32) (int, error) {
	rc := C.clk_dns_gc(e.h, C.uint32_t(now))
	if rc < 0 {
		return 0, e.err(rc)
	}
	return int(rc), nil
}

// Rule is config.EgressRule restricted to the fields on this path.
type PathRule struct{ Path, Action string }

type Rule struct {
	Dst, Proto, Action string
	Port               int
	PathRules          []PathRule // proto http: Host + path-prefix stage
	PathDefault        string
}

// RulesCompile replaces the zone table, the SNI table, the CIDR set and route_map from rules.
func (e *Engine) RulesCompile(rules []Rule, egressPort, tcpPortBase int) (warnings uint32, err error) {
	cr := make([]C.clk_rule, len(rules))
	var frees []unsafe.Pointer
	defer func() {
		for _, p := range frees {
			C.free(p)
		}
	}()
	cs := func(s string) *C.char { p := C.CString(s); frees = append(frees, unsafe.Pointer(p)); return p }
	for i, r := range rules {
		cr[i] = C.clk_rule{dst: cs(r.Dst), proto: cs(r.Proto), port: C.int32_t(r.Port), action: cs(r.Action), path_default: cs(r.PathDefault)}
		if n := len(r.PathRules); n > 0 { // the array lives in C memory: cgo forbids Go pointers to Go pointers
			arr := (*C.clk_path_rule)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.clk_path_rule{}))))
			frees = append(frees, unsafe.Pointer(arr))
			for k, pr := range r.PathRules {
				(*[1 << 20]C.clk_path_rule)(unsafe.Pointer(arr))[k] = C.clk_path_rule{path: cs(pr.Path), action: cs(pr.Action)}
			}
			cr[i].path_rules, cr[i].n_path_rules = arr, C.uint32_t(n)
		}
	}
	ports := C.clk_ports{egress_port: C.int32_t(egressPort), tcp_port_base: C.int32_t(tcpPortBase)}
	var p *C.clk_rule
	if len(cr) > 0 {
		p = &cr[0]
	}
	var w C.uint32_t
	err = e.err(C.clk_rules_compile(e.h, p, C.uint32_t(len(cr)), &ports, &w))
	return uint32(w), err
}

// Publish flips the table epoch: classification sees every change made so far.
func (e *Engine) Publish() error { return e.err(C.clk_publish(e.h)) }

// RulesDump returns one compiled table as lines (clk_rules_dump): kind 0 zones, 1 server_names, 2 TCP mappings
// "dst:port:envoy_port", 3 HTTP virtual hosts.
func (e *Engine) RulesDump(kind int) ([]string, error) {
	n := C.clk_rules_dump(e.h, C.int(kind), nil, 0)
	if n < 0 {
		return nil, e.err(n)
	}
	buf := (*C.char)(C.malloc(C.size_t(n)))
	defer C.free(unsafe.Pointer(buf))
	if rc := C.clk_rules_dump(e.h, C.int(kind), buf, C.size_t(n)); rc < 0 {
		return nil, e.err(rc)
	}
	var out []string
	for _, l := range strings.Split(C.GoString(buf), "\n") {
		if l != "" {
			out = append(out, l)
		}
	}
	return out, nil
}

// ClassifySlots classifies n = len(pkts)/stride frames laid out as fixed slots; verdicts must hold n words.
func (e *Engine) ClassifySlots(pkts []byte, stride int, flags uint32, verdicts []uint32) error {
	n := len(pkts) / stride
	if n == 0 {
		return nil
	}
	if len(verdicts) < n {
		return errors.New("clawker_gpu: verdicts too short")
	}
	return e.err(C.clk_classify_slots(e.h, (*C.uint8_t)(unsafe.Pointer(&pkts[0])), C.uint32_t(stride), C.uint64_t(n),
		C.uint32_t(flags), (*C.uint32_t)(unsafe.Pointer(&verdicts[0]))))
}

// ClassifyVarlen classifies len(offs16)-1 frames; frame i is bytes [16*offs16[i], 16*offs16[i+1]).
func (e *Engine) ClassifyVarlen(pkts []byte, offs16 []uint32, flags uint32, verdicts []uint32) error {
	n := len(offs16) - 1
	if n <= 0 {
		return nil
	}
	if len(verdicts) < n {
		return errors.New("clawker_gpu: verdicts too short")
	}
	return e.err(C.clk_classify_varlen(e.h, (*C.uint8_t)(unsafe.Pointer(&pkts[0])), (*C.uint32_t)(unsafe.Pointer(&offs16[0])),
		C.uint64_t(n), C.uint32_t(flags), (*C.uint32_t)(unsafe.Pointer(&verdicts[0]))))
}

// Metric is clk_metric: one metrics_map row summed over the GPU's counters.
type Metric struct {
	CgroupID   uint64
	DomainHash uint32
	DstPort    uint16
	Action     uint8
	_          uint8
	Count      uint64
}

// CountersRead returns this GPU's metrics_map rows.
func (e *Engine) CountersRead() ([]Metric, error) {
	var n C.uint32_t
	if err := e.err(C.clk_counters_read(e.h, nil, 0, &n)); err != nil || n == 0 {
		return nil, err
	}
	out := make([]Metric, int(n))
	err := e.err(C.clk_counters_read(e.h, (*C.clk_metric)(unsafe.Pointer(&out[0])), n, &n))
	return out[:int(n)], err
}

// Total mirrors clk_total: one container's ALLOW / DENY / BYPASS sums.
type Total struct {
	CgroupID uint64
	Count    [3]uint64
}

// TotalsRows is the fixed row count of a totals block (max_containers).
func (e *Engine) TotalsRows() (uint32, error) {
	var n C.uint32_t
	err := e.err(C.clk_totals_rows(e.h, &n))
	return uint32(n), err
}

// TotalsExport writes this GPU's per-container totals into a device block (TotalsRows x 32 B) on `stream`;
// the caller all-gathers the blocks (ncclAllGather) and hands them to TotalsMerge.
func (e *Engine) TotalsExport(dBlock, stream unsafe.Pointer) error {
	return e.err(C.clk_totals_export(e.h, dBlock, stream))
}

// TotalsMerge sums nBlocks gathered device blocks by cgroup id.
func (e *Engine) TotalsMerge(dBlocks unsafe.Pointer, nBlocks uint32) ([]Total, error) {
	rows, err := e.TotalsRows()
	if err != nil {
		return nil, err
	}
	out := make([]Total, int(rows*nBlocks))
	var n C.uint32_t
	if len(out) == 0 {
		return nil, nil
	}
	err = e.err(C.clk_totals_merge(e.h, dBlocks, C.uint32_t(nBlocks), (*C.clk_total)(unsafe.Pointer(&out[0])), C.uint32_t(len(out)), &n))
	return out[:int(n)], err
}

// StateSave / StateLoad: the pinned state shared with the ebpf-manager verbs (the bpffs analogue).
// Hdr16 is clk_hdr16: the 16-byte compact ingest record (what the decision reads of an IPv4 header-only frame).
type Hdr16 struct {
	Src, Dst     uint32 // network order bytes loaded as native u32
	Sport, Dport uint16 // host order
	Proto, Kind  uint8  // Kind: 0 IPv4 record, 1 no socket event, 2 Src indexes the 64-byte slot of the frame in aux
	_            uint16
}

// PackHdr16 turns slot-layout frames into records plus the whole frames a record cannot express (clk_pack_hdr16).
func PackHdr16(pkts []byte, stride int) (recs []Hdr16, aux []byte, err error) {
	n := len(pkts) / stride
	recs = make([]Hdr16, n)
	aux = make([]byte, 64*(n/8+1024))
	for {
		var nAux C.uint64_t
		rc := C.clk_pack_hdr16((*C.uint8_t)(unsafe.Pointer(&pkts[0])), C.uint32_t(stride), C.uint64_t(n),
			(*C.clk_hdr16)(unsafe.Pointer(&recs[0])), (*C.uint8_t)(unsafe.Pointer(&aux[0])), C.uint64_t(len(aux)/64), &nAux)
		if rc == C.CLK_OK {
			return recs, aux[:64*int(nAux)], nil
		}
		if rc != C.CLK_E2BIG || len(aux) >= 64*n {
			return nil, nil, fmt.Errorf("clk_pack_hdr16: %s", C.GoString(C.clk_last_error(nil)))
		}
		aux = make([]byte, 64*n)
	}
}

// ClassifyHdr16 classifies compact records (host buffers; 20 bytes per packet cross the host link instead of 68).
func (e *Engine) ClassifyHdr16(recs []Hdr16, aux []byte, flags uint32, verdicts []uint32) error {
	var ap *C.uint8_t
	if len(aux) > 0 {
		ap = (*C.uint8_t)(unsafe.Pointer(&aux[0]))
	}
	return e.err(C.clk_classify_hdr16(e.h, (*C.clk_hdr16)(unsafe.Pointer(&recs[0])), C.uint64_t(len(recs)), ap,
		C.uint64_t(len(aux)/64), C.uint32_t(flags), (*C.uint32_t)(unsafe.Pointer(&verdicts[0]))))
}

// CommUniqueID is ncclGetUniqueId through the library: rank 0 calls it and hands the bytes to the other ranks.
func CommUniqueID() ([]byte, error) {
	buf := make([]byte, 256)
	var n C.size_t
	if rc := C.clk_comm_unique_id(unsafe.Pointer(&buf[0]), C.size_t(len(buf)), &n); rc != C.CLK_OK {
		return nil, fmt.Errorf("clk_comm_unique_id: %s", C.GoString(C.clk_last_error(nil)))
	}
	return buf[:int(n)], nil
}

// CommInit joins this engine (one per GPU) to the host's communicator; AllgatherCounters is the per-batch collective.
func (e *Engine) CommInit(id []byte, rank, world int) error {
	return e.err(C.clk_comm_init(e.h, unsafe.Pointer(&id[0]), C.size_t(len(id)), C.int(rank), C.int(world)))
}
func (e *Engine) CommDestroy() error { return e.err(C.clk_comm_destroy(e.h)) }
func (e *Engine) AllgatherCounters(stream unsafe.Pointer) error {
	return e.err(C.clk_allgather_counters(e.h, stream))
}
func (e *Engine) AllgatherMetrics(stream unsafe.Pointer) error {
	return e.err(C.clk_allgather_metrics(e.h, stream))
}

// TotalsGlobal returns the gathered per-container totals of all GPUs, summed by cgroup id.
func (e *Engine) TotalsGlobal() ([]Total, error) {
	rows, err := e.TotalsRows()
	if err != nil {
		return nil, err
	}
	var world C.int
	if rc := C.clk_comm_info(e.h, nil, &world); rc != C.CLK_OK {
		return nil, e.err(rc)
	}
	out := make([]Total, int(rows)*int(world))
	var n C.uint32_t
	if rc := C.clk_totals_read_global(e.h, (*C.clk_total)(unsafe.Pointer(&out[0])), C.uint32_t(len(out)), &n); rc != C.CLK_OK {
		return nil, e.err(rc)
	}
	return out[:int(n)], nil
}

func (e *Engine) StateSave(path string) error {
	p := C.CString(path)
	defer C.free(unsafe.Pointer(p))
	return e.err(C.clk_state_save(e.h, p))
}

func (e *Engine) StateLoad(path string) error {
	p := C.CString(path)
	defer C.free(unsafe.Pointer(p))
	return e.err(C.clk_state_load(e.h, p))
}

// ClassifyPcap classifies every frame of a classic pcap file (read, packed and classified in overlapped chunks).
func (e *Engine) ClassifyPcap(path string, flags uint32) ([]uint32, error) {
	p := C.CString(path)
	defer C.free(unsafe.Pointer(p))
	var n C.uint64_t
	if rc := C.clk_classify_pcap(e.h, p, C.uint32_t(flags), nil, 0, &n); rc != C.CLK_OK {
		return nil, fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_pcap_last_error()), int(rc))
	}
	out := make([]uint32, int(n))
	if n == 0 {
		return out, nil
	}
	if rc := C.clk_classify_pcap(e.h, p, C.uint32_t(flags), (*C.uint32_t)(unsafe.Pointer(&out[0])), n, &n); rc != C.CLK_OK {
		return nil, fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_pcap_last_error()), int(rc))
	}
	return out[:int(n)], nil
}

// Ring is an AF_PACKET / TPACKET_V3 receive ring on one interface (clk_afpacket_*): live ingest into the varlen layout.
type Ring struct{ h *C.clk_ring }

func OpenRing(ifname string, ringMiB uint32) (*Ring, error) {
	p := C.CString(ifname)
	defer C.free(unsafe.Pointer(p))
	var h *C.clk_ring
	if rc := C.clk_afpacket_open(p, C.uint32_t(ringMiB), &h); rc != C.CLK_OK {
		return nil, fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_ingest_last_error()), int(rc))
	}
	return &Ring{h}, nil
}

// Read drains ready blocks into blob / offs16 (len(offs16) - 1 frames at most) and returns the number of frames.
func (r *Ring) Read(blob []byte, offs16 []uint32, timeoutMs int) (int, error) {
	if len(blob) == 0 || len(offs16) < 2 {
		return 0, fmt.Errorf("clawker_gpu: empty buffers")
	}
	var n C.uint64_t
	rc := C.clk_afpacket_read(r.h, (*C.uint8_t)(unsafe.Pointer(&blob[0])), C.uint64_t(len(blob)), (*C.uint32_t)(unsafe.Pointer(&offs16[0])),
		C.uint64_t(len(offs16)-1), C.int(timeoutMs), &n)
	if rc != C.CLK_OK {
		return 0, fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_ingest_last_error()), int(rc))
	}
	return int(n), nil
}

// ReadHdr16 drains ready blocks straight into compact ingest records (+ 64-byte aux slots): input of ClassifyHdr16.
func (r *Ring) ReadHdr16(recs []Hdr16, aux []byte, timeoutMs int) (n, nAux int, err error) {
	if len(recs) == 0 {
		return 0, 0, fmt.Errorf("clawker_gpu: empty buffers")
	}
	var cn, ca C.uint64_t
	var pa *C.uint8_t
	if len(aux) > 0 {
		pa = (*C.uint8_t)(unsafe.Pointer(&aux[0]))
	}
	rc := C.clk_afpacket_read_hdr16(r.h, (*C.clk_hdr16)(unsafe.Pointer(&recs[0])), C.uint64_t(len(recs)), pa, C.uint64_t(len(aux)/64),
		C.int(timeoutMs), &cn, &ca)
	if rc != C.CLK_OK {
		return 0, 0, fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_ingest_last_error()), int(rc))
	}
	return int(cn), int(ca), nil
}

func (r *Ring) Close() { C.clk_afpacket_close(r.h); r.h = nil }

// TCPCoalesce folds the continuation segments of split ClientHellos / HTTP request heads into their first segment
// (one output frame per input frame); out must hold 2 x len(blob).
func TCPCoalesce(blob []byte, offs16 []uint32, out []byte, outOffs16 []uint32) (merged uint64, err error) {
	n := len(offs16) - 1
	if n <= 0 || len(out) == 0 || len(outOffs16) < len(offs16) {
		return 0, fmt.Errorf("clawker_gpu: bad buffers")
	}
	var m C.uint64_t
	rc := C.clk_tcp_coalesce((*C.uint8_t)(unsafe.Pointer(&blob[0])), (*C.uint32_t)(unsafe.Pointer(&offs16[0])), C.uint64_t(n),
		(*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint64_t(len(out)), (*C.uint32_t)(unsafe.Pointer(&outOffs16[0])), &m)
	if rc != C.CLK_OK {
		return 0, fmt.Errorf("clawker_gpu: %s (%d)", C.GoString(C.clk_ingest_last_error()), int(rc))
	}
	return uint64(m), nil
}
