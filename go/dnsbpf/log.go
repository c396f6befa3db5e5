package dnsbpf

import clog "github.com/coredns/coredns/plugin/pkg/log"

var log = clog.NewWithPlugin(pluginName)
