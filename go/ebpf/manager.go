// Package ebpf is the drop-in for github.com/schmitthub/clawker/internal/ebpf backed by the GPU engine.
//
// Exported names, signatures and error behaviour follow internal/ebpf/manager.go and types.go so that
// internal/firewall.Manager and internal/ebpf/cmd keep compiling unchanged.  Attaching BPF programs to a
// cgroup has no counterpart: a cgroup is enforced while it has a container_map row, and traffic reaches the
// classifier through ClassifyBatch (fed by the capture/ingest side) instead of through kernel hooks.
package ebpf

import (
	"encoding/binary"
	"errors"
	"fmt"
	"net"
	"strings"
	"unsafe"

	"github.com/schmitthub/clawker/go/clkgpu"
)

// PinPath is kept for interface parity; nothing is pinned.
const PinPath = "/sys/fs/bpf/clawker"

// Row layouts: byte-identical to internal/ebpf/bpf/common.h (and to include/clawker_gpu.h).
type ContainerConfig struct {
	EnvoyIP, CoreDNSIP, GatewayIP, NetAddr, NetMask, HostProxyIP uint32
	HostProxyPort, EgressPort                                   uint16
}

type DNSEntry struct{ DomainHash, ExpireTS uint32 }

type RouteKey struct {
	DomainHash uint32
	DstPort    uint16
	_          uint16
}

type RouteVal struct {
	EnvoyPort uint16
	_         uint16
}

type MetricKey struct {
	CgroupID   uint64
	DomainHash uint32
	DstPort    uint16
	Action     uint8
	_          uint8
}

const (
	ActionAllow  uint8 = 0
	ActionDeny   uint8 = 1
	ActionBypass uint8 = 2
)

// Route describes a per-domain TCP route (JSON shape of `ebpf-manager sync-routes`).
type Route struct {
	DomainHash uint32 `json:"domain_hash"`
	DstPort    uint16 `json:"dst_port"`
	EnvoyPort  uint16 `json:"envoy_port"`
}

// ErrKeyNotExist is returned by LookupContainer for an unknown cgroup.
var ErrKeyNotExist = clkgpu.ErrKeyNotExist

func IPToUint32(ip net.IP) uint32 {
	ip = ip.To4()
	if ip == nil {
		return 0
	}
	return binary.NativeEndian.Uint32(ip)
}

func Uint32ToIP(n uint32) net.IP {
	ip := make(net.IP, 4)
	binary.NativeEndian.PutUint32(ip, n)
	return ip
}

func CIDRToAddrMask(cidr string) (addr, mask uint32, err error) {
	_, ipNet, err := net.ParseCIDR(cidr)
	if err != nil {
		return 0, 0, err
	}
	return binary.NativeEndian.Uint32(ipNet.IP.To4()), binary.NativeEndian.Uint32(net.IP(ipNet.Mask).To4()), nil
}

// DomainHash is FNV-1a-32 of the lower-cased name.
func DomainHash(domain string) uint32 {
	h := uint32(0x811c9dc5)
	for _, c := range []byte(strings.ToLower(domain)) {
		h = (h ^ uint32(c)) * 0x01000193
	}
	return h
}

func NewContainerConfig(envoyIP, corednsIP, gatewayIP, cidr, hostProxyIP string, hostProxyPort, egressPort uint16) (ContainerConfig, error) {
	netAddr, netMask, err := CIDRToAddrMask(cidr)
	if err != nil {
		return ContainerConfig{}, fmt.Errorf("parsing CIDR %s: %w", cidr, err)
	}
	cfg := ContainerConfig{EnvoyIP: IPToUint32(net.ParseIP(envoyIP)), CoreDNSIP: IPToUint32(net.ParseIP(corednsIP)),
		GatewayIP: IPToUint32(net.ParseIP(gatewayIP)), NetAddr: netAddr, NetMask: netMask, HostProxyPort: hostProxyPort, EgressPort: egressPort}
	if hostProxyIP != "" {
		cfg.HostProxyIP = IPToUint32(net.ParseIP(hostProxyIP))
	}
	return cfg, nil
}

// Manager owns one GPU engine.
type Manager struct {
	eng *clkgpu.Engine
	cfg clkgpu.Config
}

func NewManager(cfg clkgpu.Config) *Manager { return &Manager{cfg: cfg} }

// Load creates the tables (the reference loads and pins BPF objects here).
func (m *Manager) Load() error {
	if m.eng != nil {
		return nil
	}
	e, err := clkgpu.New(m.cfg)
	if err != nil {
		return fmt.Errorf("ebpf load: %w", err)
	}
	m.eng = e
	return nil
}

// OpenPinned is Load: there is no cross-process pin, the engine lives in this process.
func (m *Manager) OpenPinned() error { return m.Load() }

func (m *Manager) Close() error {
	if m.eng == nil {
		return nil
	}
	err := m.eng.Close()
	m.eng = nil
	return err
}

func (m *Manager) clearBypass(cgroupID uint64) error {
	err := m.eng.MapDelete(clkgpu.MapBypass, unsafe.Pointer(&cgroupID), 8)
	if err == nil || errors.Is(err, clkgpu.ErrKeyNotExist) {
		return nil
	}
	return err
}

// Enable clears a stale bypass flag and installs the container's config; the row becomes visible at Publish.
func (m *Manager) Enable(cgroupID uint64, cgroupPath string, cfg ContainerConfig) error {
	_ = m.clearBypass(cgroupID)
	if err := m.eng.MapUpdate(clkgpu.MapContainer, unsafe.Pointer(&cgroupID), 8, unsafe.Pointer(&cfg), unsafe.Sizeof(cfg)); err != nil {
		return fmt.Errorf("ebpf enable: updating container_map: %w", err)
	}
	return m.eng.Publish()
}

func (m *Manager) Disable(cgroupID uint64) error {
	if err := m.eng.MapDelete(clkgpu.MapContainer, unsafe.Pointer(&cgroupID), 8); err != nil && !errors.Is(err, clkgpu.ErrKeyNotExist) {
		return err
	}
	_ = m.clearBypass(cgroupID)
	return m.eng.Publish()
}

func (m *Manager) SyncRoutes(routes []Route) error {
	rs := make([]clkgpu.Route, len(routes))
	for i, r := range routes {
		rs[i] = clkgpu.Route{DomainHash: r.DomainHash, DstPort: r.DstPort, EnvoyPort: r.EnvoyPort}
	}
	if err := m.eng.RoutesReplace(rs); err != nil {
		return err
	}
	return m.eng.Publish()
}

func (m *Manager) Bypass(cgroupID uint64) error {
	one := uint8(1)
	if err := m.eng.MapUpdate(clkgpu.MapBypass, unsafe.Pointer(&cgroupID), 8, unsafe.Pointer(&one), 1); err != nil {
		return fmt.Errorf("ebpf bypass: %w", err)
	}
	return m.eng.Publish()
}

func (m *Manager) Unbypass(cgroupID uint64) error {
	if err := m.clearBypass(cgroupID); err != nil {
		return fmt.Errorf("ebpf unbypass: %w", err)
	}
	return m.eng.Publish()
}

func (m *Manager) UpdateDNSCache(ip, domainHash, ttlSeconds uint32) error {
	return m.eng.DNSUpdate(ip, domainHash, ttlSeconds, 0)
}

func (m *Manager) GarbageCollectDNS() int {
	n, err := m.eng.DNSGC(0)
	if err != nil {
		return 0
	}
	return n
}

func (m *Manager) LookupContainer(cgroupID uint64) (ContainerConfig, error) {
	var cfg ContainerConfig
	err := m.eng.MapLookup(clkgpu.MapContainer, unsafe.Pointer(&cgroupID), 8, unsafe.Pointer(&cfg), unsafe.Sizeof(cfg))
	return cfg, err
}

// ClassifyBatch is the new verb: one verdict word per 64-byte slot (see include/clawker_gpu.h for the layout).
func (m *Manager) ClassifyBatch(pkts []byte, stride int, verdicts []uint32) error {
	return m.eng.ClassifySlots(pkts, stride, 0, verdicts)
}

// Engine exposes the underlying engine to the dnsbpf plugin and the firewall rule compiler.
func (m *Manager) Engine() *clkgpu.Engine { return m.eng }
