// Package dnsbpf is the drop-in for github.com/schmitthub/clawker/internal/dnsbpf: the CoreDNS plugin that
// records resolved A records, writing into the GPU engine's dns_cache instead of a pinned BPF map.
// Plugin name, directive placement and Handler behaviour are the reference's (dnsbpf.go:32-105, setup.go:12-65).
package dnsbpf

import (
	"context"
	"strings"

	"github.com/coredns/caddy"
	"github.com/coredns/coredns/core/dnsserver"
	"github.com/coredns/coredns/plugin"
	corelog "github.com/coredns/coredns/plugin/pkg/log"
	"github.com/coredns/coredns/plugin/pkg/nonwriter"
	"github.com/miekg/dns"

	clawkerebpf "github.com/schmitthub/clawker/go/ebpf"
)

const pluginName = "dnsbpf"

var warn = corelog.NewWithPlugin(pluginName).Warningf // cache-write failures are logged, never fatal to the query

// MapWriter is the interface for writing dns_cache rows.
type MapWriter interface {
	Update(ip, domainHash, ttlSeconds uint32)
}

// EngineMap writes through a shared ebpf.Manager (engine handle).
type EngineMap struct{ M *clawkerebpf.Manager }

func (b EngineMap) Update(ip, domainHash, ttlSeconds uint32) {
	if err := b.M.UpdateDNSCache(ip, domainHash, ttlSeconds); err != nil {
		warn("updating dns_cache for ip=%s hash=%d: %v", clawkerebpf.Uint32ToIP(ip), domainHash, err)
	}
}

// Handler intercepts responses of the next plugin and records their A records.
type Handler struct {
	Next plugin.Handler
	Zone string
	Map  MapWriter
}

func (h Handler) Name() string { return pluginName }

func (h Handler) ServeDNS(ctx context.Context, w dns.ResponseWriter, r *dns.Msg) (int, error) {
	nw := nonwriter.New(w)
	rcode, err := plugin.NextOrFailure(h.Name(), h.Next, ctx, nw, r)
	if err != nil || nw.Msg == nil {
		return rcode, err
	}
	if h.Map != nil {
		hash := clawkerebpf.DomainHash(strings.TrimSuffix(h.Zone, "."))
		for _, rr := range nw.Msg.Answer {
			a, ok := rr.(*dns.A)
			if !ok {
				continue
			}
			ip := clawkerebpf.IPToUint32(a.A)
			if ip == 0 {
				continue
			}
			ttl := rr.Header().Ttl
			if ttl == 0 {
				ttl = 60
			}
			h.Map.Update(ip, hash, ttl)
		}
	}
	if err := w.WriteMsg(nw.Msg); err != nil {
		return dns.RcodeServerFailure, err
	}
	return rcode, nil
}

// Shared is set by the embedding binary (cmd/coredns-clawker) before CoreDNS starts.
var Shared MapWriter

func init() { plugin.Register(pluginName, setup) }

func setup(c *caddy.Controller) error {
	for c.Next() {
		if c.NextArg() {
			return plugin.Error(pluginName, c.ArgErr())
		}
	}
	if Shared == nil {
		return plugin.Error(pluginName, c.Err("GPU engine dns_cache is not available (is the engine created?)"))
	}
	zone := dnsserver.GetConfig(c).Zone
	dnsserver.GetConfig(c).AddPlugin(func(next plugin.Handler) plugin.Handler {
		return Handler{Next: next, Zone: zone, Map: Shared}
	})
	return nil
}
