// Package firewall holds the rule-facing entry points of the drop-in: the table-producing part of
// internal/firewall (ValidateDst, normalizeAndDedup, GenerateCorefile's zone set, serverNames, TCPMappings,
// syncRoutes) is computed by the C++ rule compiler inside libclawker_gpu.so.
package firewall

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -lclawker_gpu
#include <stdlib.h>
#include "clawker_gpu.h"
*/
import "C"

import (
	"fmt"
	"strconv"
	"strings"
	"unsafe"

	"github.com/schmitthub/clawker/go/clkgpu"
)

// EgressRule is config.EgressRule without the path-rule fields (those are not on this path).
type EgressRule struct {
	Dst    string `yaml:"dst"`
	Proto  string `yaml:"proto,omitempty"`
	Port   int    `yaml:"port,omitempty"`
	Action string `yaml:"action,omitempty"`
}

// EnvoyPorts mirrors the reference's port configuration.
type EnvoyPorts struct{ EgressPort, TCPPortBase, HealthPort int }

// ValidateDst accepts the same destination grammar as the reference.
func ValidateDst(dst string) error {
	cs := C.CString(dst)
	defer C.free(unsafe.Pointer(cs))
	if C.clk_validate_dst(cs) != C.CLK_OK {
		return fmt.Errorf("invalid destination %q", dst)
	}
	return nil
}

// CompileRules installs zone table, SNI table, CIDR set and route_map for the rule set and publishes them.
func CompileRules(e *clkgpu.Engine, rules []EgressRule, ports EnvoyPorts) (warnings uint32, err error) {
	rs := make([]clkgpu.Rule, len(rules))
	for i, r := range rules {
		rs[i] = clkgpu.Rule{Dst: r.Dst, Proto: r.Proto, Action: r.Action, Port: r.Port}
	}
	if warnings, err = e.RulesCompile(rs, ports.EgressPort, ports.TCPPortBase); err != nil {
		return warnings, err
	}
	return warnings, e.Publish()
}

// TCPMapping is firewall.TCPMapping (internal/firewall/envoy.go:40-48): one tcp/ssh allow rule and the Envoy listener port it owns.
type TCPMapping struct {
	Dst       string
	DstPort   int
	EnvoyPort int
}

// TCPMappings is the reference's firewall.TCPMappings(rules, ports) (internal/firewall/envoy.go:54-83): the mappings the
// engine derived when the rules were compiled, in rule order (EnvoyPort = TCPPortBase + index).  CompileRules must have run.
func TCPMappings(e *clkgpu.Engine) ([]TCPMapping, error) {
	lines, err := e.RulesDump(2)
	if err != nil {
		return nil, err
	}
	out := make([]TCPMapping, 0, len(lines))
	for _, l := range lines {
		f := strings.Split(l, ":")
		if len(f) != 3 {
			return nil, fmt.Errorf("malformed mapping %q", l)
		}
		dp, err1 := strconv.Atoi(f[1])
		ep, err2 := strconv.Atoi(f[2])
		if err1 != nil || err2 != nil {
			return nil, fmt.Errorf("malformed mapping %q", l)
		}
		out = append(out, TCPMapping{Dst: f[0], DstPort: dp, EnvoyPort: ep})
	}
	return out, nil
}
