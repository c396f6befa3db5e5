// Command coredns-clawker: the CoreDNS build that carries the dnsbpf plugin of this repository (go/dnsbpf), i.e. the
// registration surface of the reference's cmd/coredns-clawker/main.go:30-34 with the GPU engine behind the plugin.
//
// What has to hold for the drop-in (and is all this file does):
//   - the plugins the generated Corefile names are linked in (forward, health, log, reload, template —
//     internal/firewall/coredns.go:66-96 emits exactly these directives);
//   - "dnsbpf" is FIRST in dnsserver.Directives, so the plugin wraps every resolver plugin and sees every response
//     before it leaves (the answer filter -> clk_dns_update path);
//   - the process is coremain.Run(), so flags / Corefile handling / `reload 2s` behave as in the reference image.
//
// Not compiled in this image (no Go toolchain, see go/README.md).
package main

import (
	"github.com/coredns/coredns/core/dnsserver"
	"github.com/coredns/coredns/coremain"

	_ "github.com/coredns/coredns/plugin/forward"
	_ "github.com/coredns/coredns/plugin/health"
	_ "github.com/coredns/coredns/plugin/log"
	_ "github.com/coredns/coredns/plugin/reload"
	_ "github.com/coredns/coredns/plugin/template"

	_ "github.com/schmitthub/clawker/go/dnsbpf" // registers plugin "dnsbpf" (setup.go:12-65 equivalent)
)

func init() {
	rest := make([]string, 0, len(dnsserver.Directives)+1)
	rest = append(rest, "dnsbpf")
	for _, d := range dnsserver.Directives {
		if d != "dnsbpf" {
			rest = append(rest, d)
		}
	}
	dnsserver.Directives = rest
}

func main() { coremain.Run() }
