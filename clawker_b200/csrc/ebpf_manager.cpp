// ebpf-manager — the control-plane verbs of the reference's eBPF manager container, on the GPU engine's
// pinned state.
//
// Replaces internal/ebpf/cmd/main.go:30-303 verb for verb (same argv, same JSON payloads, same messages and
// exit codes), so the unmodified firewall.Manager can keep driving enforcement through `docker exec
// ebpf-manager <verb> ...` (internal/firewall/manager.go:504-590,1356-1377).  Where the reference opens the
// maps pinned in bpffs, this binary opens the engine's state file (clk_state_load), applies the verb through
// the C ABI on a host-only engine (no GPU needed in the control-plane container) and writes it back
// (clk_state_save); the classification process reloads the file and publishes.  Concurrent invocations are
// serialised with flock, which is what the kernel's map locks give the reference.
//
//   CLK_PIN_PATH     state file (default /run/clawker-gpu/maps.pin — the /sys/fs/bpf/clawker analogue)
//   CLK_CGROUP_ROOT  cgroup v2 root the paths must live under (default /sys/fs/cgroup/; tests point it elsewhere)
//
// Two verbs have no reference counterpart: `bind` / `unbind` maintain the container-address -> {cgroup, uid,
// SO_MARK} rows the packet form needs (the hooks read those from the calling task).  `enable` also accepts the
// optional keys container_ip / uid / so_mark in its JSON and binds in the same transaction.
#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cctype>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <vector>

#include "../../include/clawker_gpu.h"

namespace {

[[noreturn]] void fatal(const char *cmd, const char *fmt, ...)
{
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    fprintf(stderr, "ebpf-manager %s: %s\n", cmd, buf);               // main.go:300-303
    exit(1);
}

void usage()
{
    fputs("Usage: ebpf-manager <command> [args...]\n\n"
          "Commands:\n"
          "  init                                    Create the pinned state\n"
          "  enable  <cgroupPath> <configJSON>       Add to container_map\n"
          "  disable <cgroupPath>                    Remove from container_map\n"
          "  sync-routes <routesJSON>                Replace global route_map\n"
          "  bypass  <cgroupPath>                    Set bypass flag\n"
          "  unbypass <cgroupPath>                   Clear bypass flag\n"
          "  dns-update <ip> <domainHash> <ttl>      Update DNS cache entry\n"
          "  gc-dns                                  Remove expired DNS cache entries\n"
          "  dump <cgroupPath>                       Print the container_map entry\n"
          "  resolve <hostname>                      First IPv4 address of a host\n"
          "  bind <cgroupPath> <ip> [uid] [so_mark]  Bind a container address to its cgroup (packet form)\n"
          "  unbind <ip>                             Remove an address binding\n", stderr);
}

// ---- a JSON reader for the two payload shapes of main.go:30-38 and manager.go:466-470 -----------------------
struct JVal { enum { STR, NUM, OTHER } kind = OTHER; std::string s; };
typedef std::map<std::string, JVal> JObj;

struct JParser {
    const char *p; std::string err;
    explicit JParser(const char *s) : p(s) {}
    void ws() { while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++; }
    bool fail(const char *m) { if (err.empty()) err = m; return false; }
    bool string(std::string &out)
    {
        if (*p != '"') return fail("invalid character looking for beginning of string");
        p++;
        while (*p && *p != '"') {
            if (*p == '\\') {
                p++;
                switch (*p) {
                case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
                case 'r': out += '\r'; break; case 't': out += '\t'; break;
                case 'u': {
                    unsigned v = 0;
                    for (int i = 1; i <= 4; i++) { if (!isxdigit((unsigned char)p[i])) return fail("invalid \\u escape"); v = v * 16 + (isdigit((unsigned char)p[i]) ? p[i] - '0' : (tolower(p[i]) - 'a' + 10)); }
                    p += 4;
                    if (v < 0x80) out += (char)v; else if (v < 0x800) { out += (char)(0xC0 | (v >> 6)); out += (char)(0x80 | (v & 63)); }
                    else { out += (char)(0xE0 | (v >> 12)); out += (char)(0x80 | ((v >> 6) & 63)); out += (char)(0x80 | (v & 63)); }
                    break;
                }
                default: return fail("invalid escape in string");
                }
                p++;
            } else out += *p++;
        }
        if (*p != '"') return fail("unexpected end of JSON input");
        p++;
        return true;
    }
    bool value(JVal &v)
    {
        ws();
        if (*p == '"') { v.kind = JVal::STR; return string(v.s); }
        if (*p == '-' || isdigit((unsigned char)*p)) {
            v.kind = JVal::NUM; const char *b = p;
            if (*p == '-') p++;
            while (isdigit((unsigned char)*p) || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-') p++;
            v.s.assign(b, p - b);
            return true;
        }
        for (const char *lit : { "true", "false", "null" }) if (!strncmp(p, lit, strlen(lit))) { v.kind = JVal::OTHER; v.s = lit; p += strlen(lit); return true; }
        return fail("invalid character looking for beginning of value");
    }
    bool object(JObj &o)
    {
        ws();
        if (*p != '{') return fail("cannot unmarshal: expected an object");
        p++; ws();
        if (*p == '}') { p++; return true; }
        for (;;) {
            ws(); std::string k;
            if (!string(k)) return false;
            for (auto &c : k) c = (char)tolower((unsigned char)c);      // encoding/json matches keys case-insensitively
            ws(); if (*p != ':') return fail("invalid character after object key");
            p++;
            JVal v; if (!value(v)) return false;
            o[k] = v; ws();
            if (*p == ',') { p++; continue; }
            if (*p == '}') { p++; return true; }
            return fail("invalid character after object key:value pair");
        }
    }
    bool array(std::vector<JObj> &a)
    {
        ws();
        if (!strncmp(p, "null", 4)) { p += 4; return true; }
        if (*p != '[') return fail("cannot unmarshal: expected an array");
        p++; ws();
        if (*p == ']') { p++; return true; }
        for (;;) {
            JObj o; if (!object(o)) return false;
            a.push_back(o); ws();
            if (*p == ',') { p++; continue; }
            if (*p == ']') { p++; return true; }
            return fail("invalid character after array element");
        }
    }
    bool end() { ws(); return *p == 0 || fail("invalid character after top-level value"); }
};

bool get_str(const JObj &o, const char *key, std::string &out, std::string &err)
{
    auto it = o.find(key);
    if (it == o.end() || (it->second.kind == JVal::OTHER && it->second.s == "null")) { out.clear(); return true; }
    if (it->second.kind != JVal::STR) { err = std::string("json: cannot unmarshal non-string into field ") + key; return false; }
    out = it->second.s; return true;
}
bool get_uint(const JObj &o, const char *key, uint64_t max, uint64_t &out, std::string &err)
{
    out = 0;
    auto it = o.find(key);
    if (it == o.end() || (it->second.kind == JVal::OTHER && it->second.s == "null")) return true;
    const std::string &s = it->second.s;
    if (it->second.kind != JVal::NUM || s.find_first_not_of("0123456789") != std::string::npos || s.size() > 19) { err = std::string("json: cannot unmarshal ") + s + " into unsigned field " + key; return false; }
    out = strtoull(s.c_str(), nullptr, 10);
    if (out > max) { err = std::string("json: cannot unmarshal number ") + s + " into field " + key + " (out of range)"; return false; }
    return true;
}

// ---- cgroup paths: validateCgroupPath + CgroupID (manager.go:511-555) ---------------------------------------
std::string cgroup_root()
{
    const char *r = getenv("CLK_CGROUP_ROOT");
    std::string s = (r && *r) ? r : "/sys/fs/cgroup/";
    if (s.back() != '/') s += '/';
    return s;
}
std::string path_clean(const std::string &p)                            // filepath.Clean for a path without ".."
{
    std::vector<std::string> parts; std::string cur;
    for (size_t i = 0; i <= p.size(); i++) {
        if (i == p.size() || p[i] == '/') { if (!cur.empty() && cur != ".") parts.push_back(cur); cur.clear(); }
        else cur += p[i];
    }
    std::string out = (!p.empty() && p[0] == '/') ? "/" : "";
    for (size_t i = 0; i < parts.size(); i++) { if (i) out += '/'; out += parts[i]; }
    return out.empty() ? "." : out;
}
uint64_t cgroup_id(const char *cmd, const char *arg, bool wrap)
{
    std::string p = arg;
    const char *pre = wrap ? "getting cgroup ID for " : "";
    std::string ctx = wrap ? std::string(pre) + p + ": " : "";
    if (p.empty()) fatal(cmd, "%scgroup path is empty", ctx.c_str());
    if (p.find_first_of(std::string("\0\n\r", 3)) != std::string::npos) fatal(cmd, "%scgroup path contains illegal characters: \"%s\"", ctx.c_str(), p.c_str());
    if (p.find("..") != std::string::npos) fatal(cmd, "%scgroup path must not contain '..': \"%s\"", ctx.c_str(), p.c_str());
    const std::string cleaned = path_clean(p), root = cgroup_root();
    if (cleaned.compare(0, root.size(), root) != 0) fatal(cmd, "%scgroup path must be under %s: \"%s\"", ctx.c_str(), root.c_str(), p.c_str());
    struct stat st;
    if (stat(cleaned.c_str(), &st) != 0) fatal(cmd, "%sopening cgroup: open %s: %s", ctx.c_str(), cleaned.c_str(), strerror(errno));
    return (uint64_t)st.st_ino;                                         // the cgroup id is the inode number on cgroup v2
}

// ---- the pinned state -----------------------------------------------------------------------------------------
struct Pinned {
    clk_engine *e = nullptr; std::string path; int lock_fd = -1; const char *cmd;
    Pinned(const char *cmd_, bool create) : cmd(cmd_)
    {
        const char *p = getenv("CLK_PIN_PATH");
        path = (p && *p) ? p : "/run/clawker-gpu/maps.pin";
        const std::string lock = path + ".lock";
        if (create) {
            const size_t slash = path.rfind('/');
            if (slash != std::string::npos && slash > 0) {
                std::string dir = path.substr(0, slash);
                for (size_t i = 1; i <= dir.size(); i++) if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0700);
            }
        }
        else if (access(path.c_str(), F_OK) != 0 && errno == ENOENT)
            fatal(cmd, "opening pinned maps: %s does not exist (run `ebpf-manager init` first)", path.c_str());   // OpenPinned manager.go:111-133
        lock_fd = open(lock.c_str(), O_CREAT | O_RDWR, 0600);
        if (lock_fd < 0 || flock(lock_fd, LOCK_EX) != 0) fatal(cmd, "locking %s: %s", lock.c_str(), strerror(errno));
        clk_cfg cfg; memset(&cfg, 0, sizeof cfg);
        cfg.abi_version = CLK_ABI_VERSION; cfg.device = -1; cfg.flags = CLK_CFG_HOST_ONLY;
        if (clk_engine_create(&cfg, &e) != CLK_OK) fatal(cmd, "%s", clk_last_error(nullptr));
        const int rc = clk_state_load(e, path.c_str());
        if (rc == CLK_ENOENT && create) return;                           // init on a clean host
        if (rc == CLK_ENOENT) fatal(cmd, "opening pinned maps: %s does not exist (run `ebpf-manager init` first)", path.c_str());   // OpenPinned manager.go:111-133
        if (rc != CLK_OK) fatal(cmd, "opening pinned maps: %s", clk_last_error(e));
    }
    void save() { if (clk_state_save(e, path.c_str()) != CLK_OK) fatal(cmd, "%s", clk_last_error(e)); }
    ~Pinned() { if (e) clk_engine_destroy(e); if (lock_fd >= 0) close(lock_fd); }
};

uint32_t ip_or_zero(const std::string &s) { uint32_t v = 0; if (!s.empty()) clk_ip_to_u32(s.c_str(), &v); return v; }   // IPToUint32(parseIP(s)): 0 for a bad address
std::string u32_to_ip(uint32_t v) { char b[32]; const uint8_t *p = (const uint8_t *)&v; snprintf(b, sizeof b, "%u.%u.%u.%u", p[0], p[1], p[2], p[3]); return b; }   // types.go:95-99

void bind_endpoint(Pinned &pin, const char *ipstr, uint64_t cg, uint64_t uid, uint64_t mark)
{
    clk_endpoint ep; memset(&ep, 0, sizeof ep); ep.cgroup_id = cg; ep.uid = (uint32_t)uid; ep.so_mark = (uint32_t)mark;
    uint32_t ip4;
    if (clk_ip_to_u32(ipstr, &ip4) == CLK_OK && strchr(ipstr, ':') == nullptr) {
        if (clk_map_update(pin.e, CLK_MAP_ENDPOINT, &ip4, 4, &ep, sizeof ep) != CLK_OK) fatal(pin.cmd, "binding %s: %s", ipstr, clk_last_error(pin.e));
        return;
    }
    uint8_t a6[16];
    if (inet_pton(AF_INET6, ipstr, a6) != 1) fatal(pin.cmd, "invalid IP: %s", ipstr);
    if (clk_map_update(pin.e, CLK_MAP_ENDPOINT6, a6, 16, &ep, sizeof ep) != CLK_OK) fatal(pin.cmd, "binding %s: %s", ipstr, clk_last_error(pin.e));
}

int ok_or_missing(int rc) { return rc == CLK_ENOENT ? CLK_OK : rc; }     // ErrKeyNotExist is success (manager.go:222-231,301-337)

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 2) { usage(); return 1; }
    const std::string cmd = argv[1];
    auto need = [&](int n) { if (argc < n) { usage(); exit(1); } };

    if (cmd == "init") {                                                  // runInit main.go:84-92
        Pinned pin("init", true);
        pin.save();
        puts("eBPF programs loaded and pinned");
        return 0;
    }
    if (cmd == "enable") {                                                // runEnable main.go:94-124
        need(4);
        JParser jp(argv[3]); JObj o;
        if (!jp.object(o) || !jp.end()) fatal("enable", "parsing config JSON: %s", jp.err.c_str());
        std::string envoy, coredns, gateway, cidr, hp_ip, cip, err; uint64_t hp_port = 0, egress = 0, uid = 1000, mark = 0;
        if (!get_str(o, "envoy_ip", envoy, err) || !get_str(o, "coredns_ip", coredns, err) || !get_str(o, "gateway_ip", gateway, err) ||
            !get_str(o, "cidr", cidr, err) || !get_str(o, "host_proxy_ip", hp_ip, err) || !get_uint(o, "host_proxy_port", 65535, hp_port, err) ||
            !get_uint(o, "egress_port", 65535, egress, err) || !get_str(o, "container_ip", cip, err))
            fatal("enable", "parsing config JSON: %s", err.c_str());
        if (o.count("uid") && !get_uint(o, "uid", 0xffffffffull, uid, err)) fatal("enable", "parsing config JSON: %s", err.c_str());
        if (o.count("so_mark") && !get_uint(o, "so_mark", 0xffffffffull, mark, err)) fatal("enable", "parsing config JSON: %s", err.c_str());
        clk_container_config cfg; memset(&cfg, 0, sizeof cfg);              // NewContainerConfig manager.go:473-494
        if (clk_cidr_to_addr_mask(cidr.c_str(), &cfg.net_addr, &cfg.net_mask) != CLK_OK) fatal("enable", "parsing CIDR %s: invalid CIDR address: %s", cidr.c_str(), cidr.c_str());
        cfg.envoy_ip = ip_or_zero(envoy); cfg.coredns_ip = ip_or_zero(coredns); cfg.gateway_ip = ip_or_zero(gateway);
        cfg.host_proxy_ip = ip_or_zero(hp_ip); cfg.host_proxy_port = (uint16_t)hp_port; cfg.egress_port = (uint16_t)egress;
        const uint64_t cg = cgroup_id("enable", argv[2], true);
        Pinned pin("enable", false);
        if (ok_or_missing(clk_map_delete(pin.e, CLK_MAP_BYPASS, &cg, 8)) != CLK_OK ||                       // Manager.Enable manager.go:236-298
            clk_map_update(pin.e, CLK_MAP_CONTAINER, &cg, 8, &cfg, sizeof cfg) != CLK_OK)
            fatal("enable", "ebpf enable: updating container_map: %s", clk_last_error(pin.e));
        if (!cip.empty()) bind_endpoint(pin, cip.c_str(), cg, uid, mark);
        pin.save();
        printf("enabled cgroup_id=%llu\n", (unsigned long long)cg);
        return 0;
    }
    if (cmd == "disable") {                                               // runDisable main.go:145-161
        need(3);
        const uint64_t cg = cgroup_id("disable", argv[2], true);
        Pinned pin("disable", false);
        if (ok_or_missing(clk_map_delete(pin.e, CLK_MAP_CONTAINER, &cg, 8)) != CLK_OK || ok_or_missing(clk_map_delete(pin.e, CLK_MAP_BYPASS, &cg, 8)) != CLK_OK)
            fatal("disable", "%s", clk_last_error(pin.e));
        pin.save();
        printf("disabled cgroup_id=%llu\n", (unsigned long long)cg);
        return 0;
    }
    if (cmd == "bypass" || cmd == "unbypass") {                           // main.go:163-197
        need(3);
        const uint64_t cg = cgroup_id(cmd.c_str(), argv[2], false);
        Pinned pin(cmd.c_str(), false);
        const uint8_t one = 1;
        const int rc = cmd == "bypass" ? clk_map_update(pin.e, CLK_MAP_BYPASS, &cg, 8, &one, 1) : ok_or_missing(clk_map_delete(pin.e, CLK_MAP_BYPASS, &cg, 8));
        if (rc != CLK_OK) fatal(cmd.c_str(), "%s", clk_last_error(pin.e));
        pin.save();
        printf("bypass %s cgroup_id=%llu\n", cmd == "bypass" ? "enabled" : "disabled", (unsigned long long)cg);
        return 0;
    }
    if (cmd == "sync-routes") {                                           // runSyncRoutes main.go:126-143
        need(3);
        JParser jp(argv[2]); std::vector<JObj> arr;
        if (!jp.array(arr) || !jp.end()) fatal("sync-routes", "parsing routes JSON: %s", jp.err.c_str());
        std::vector<clk_route> routes;
        for (const JObj &o : arr) {
            uint64_t h = 0, dp = 0, ep = 0; std::string err;
            if (!get_uint(o, "domain_hash", 0xffffffffull, h, err) || !get_uint(o, "dst_port", 65535, dp, err) || !get_uint(o, "envoy_port", 65535, ep, err))
                fatal("sync-routes", "parsing routes JSON: %s", err.c_str());
            clk_route r; memset(&r, 0, sizeof r); r.domain_hash = (uint32_t)h; r.dst_port = (uint16_t)dp; r.envoy_port = (uint16_t)ep;
            routes.push_back(r);
        }
        Pinned pin("sync-routes", false);
        if (clk_routes_replace(pin.e, routes.data(), (uint32_t)routes.size()) != CLK_OK) fatal("sync-routes", "%s", clk_last_error(pin.e));
        pin.save();
        printf("synced %zu routes\n", routes.size());
        return 0;
    }
    if (cmd == "dns-update") {                                            // runDNSUpdate main.go:199-224
        need(5);
        uint32_t ip = 0; uint8_t a6[16];
        const bool v4 = clk_ip_to_u32(argv[2], &ip) == CLK_OK;
        if (!v4 && inet_pton(AF_INET6, argv[2], a6) != 1) fatal("dns-update", "invalid IP: %s", argv[2]);
        char *end = nullptr; errno = 0;
        const unsigned long long h = strtoull(argv[3], &end, 10);
        if (errno || !*argv[3] || *end || h > 0xffffffffull || argv[3][0] == '-' || argv[3][0] == '+') fatal("dns-update", "parsing domain hash: strconv.ParseUint: parsing \"%s\": invalid syntax", argv[3]);
        errno = 0;
        const unsigned long long ttl = strtoull(argv[4], &end, 10);
        if (errno || !*argv[4] || *end || ttl > 0xffffffffull || argv[4][0] == '-' || argv[4][0] == '+') fatal("dns-update", "parsing TTL: strconv.ParseUint: parsing \"%s\": invalid syntax", argv[4]);
        Pinned pin("dns-update", false);
        if (clk_dns_update(pin.e, ip, (uint32_t)h, (uint32_t)ttl, 0) != CLK_OK) fatal("dns-update", "%s", clk_last_error(pin.e));   // a non-IPv4 address maps to 0, as IPToUint32 does
        pin.save();
        return 0;
    }
    if (cmd == "gc-dns") {                                                // runGCDNS main.go:226-235
        Pinned pin("gc-dns", false);
        const int removed = clk_dns_gc(pin.e, 0);
        if (removed < 0) fatal("gc-dns", "%s", clk_last_error(pin.e));
        pin.save();
        printf("gc-dns: removed %d expired entries\n", removed);
        return 0;
    }
    if (cmd == "dump") {                                                  // runDump main.go:237-266
        need(3);
        const uint64_t cg = cgroup_id("dump", argv[2], true);
        Pinned pin("dump", false);
        clk_container_config c;
        const int rc = clk_map_lookup(pin.e, CLK_MAP_CONTAINER, &cg, 8, &c, sizeof c);
        if (rc != CLK_OK) { printf("cgroup_id=%llu: no container_map entry: %s\n", (unsigned long long)cg, rc == CLK_ENOENT ? "key does not exist" : clk_last_error(pin.e)); return 0; }
        printf("cgroup_id=%llu\n", (unsigned long long)cg);
        printf("  envoy_ip=%s coredns_ip=%s gateway_ip=%s\n", u32_to_ip(c.envoy_ip).c_str(), u32_to_ip(c.coredns_ip).c_str(), u32_to_ip(c.gateway_ip).c_str());
        printf("  net_addr=%s net_mask=%s\n", u32_to_ip(c.net_addr).c_str(), u32_to_ip(c.net_mask).c_str());
        printf("  host_proxy_ip=%s host_proxy_port=%u egress_port=%u\n", u32_to_ip(c.host_proxy_ip).c_str(), c.host_proxy_port, c.egress_port);
        return 0;
    }
    if (cmd == "resolve") {                                               // runResolve main.go:268-280
        need(3);
        struct addrinfo hints, *res = nullptr; memset(&hints, 0, sizeof hints); hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
        const int rc = getaddrinfo(argv[2], nullptr, &hints, &res);
        if (rc != 0 || !res) fatal("resolve", "lookup %s: %s", argv[2], rc ? gai_strerror(rc) : "no IPv4 address");
        char b[64]; inet_ntop(AF_INET, &((struct sockaddr_in *)res->ai_addr)->sin_addr, b, sizeof b);
        puts(b);
        freeaddrinfo(res);
        return 0;
    }
    if (cmd == "bind") {                                                  // extension (see the header comment)
        need(4);
        const uint64_t cg = cgroup_id("bind", argv[2], true);
        Pinned pin("bind", false);
        bind_endpoint(pin, argv[3], cg, argc > 4 ? strtoull(argv[4], nullptr, 10) : 1000, argc > 5 ? strtoull(argv[5], nullptr, 0) : 0);
        pin.save();
        printf("bound %s cgroup_id=%llu\n", argv[3], (unsigned long long)cg);
        return 0;
    }
    if (cmd == "unbind") {
        need(3);
        Pinned pin("unbind", false);
        uint32_t ip4; uint8_t a6[16]; int rc;
        if (clk_ip_to_u32(argv[2], &ip4) == CLK_OK && !strchr(argv[2], ':')) rc = clk_map_delete(pin.e, CLK_MAP_ENDPOINT, &ip4, 4);
        else if (inet_pton(AF_INET6, argv[2], a6) == 1) rc = clk_map_delete(pin.e, CLK_MAP_ENDPOINT6, a6, 16);
        else fatal("unbind", "invalid IP: %s", argv[2]);
        if (ok_or_missing(rc) != CLK_OK) fatal("unbind", "%s", clk_last_error(pin.e));
        pin.save();
        printf("unbound %s\n", argv[2]);
        return 0;
    }
    fprintf(stderr, "unknown command: %s\n", cmd.c_str());
    usage();
    return 1;
}
